"""Two updates per sweep (csrc/lbm_fused.cuh, the default of lbm_step_n / lbm_compute_frames) against the
single-update kernels (LBM_FLAG_NO_FUSE) and against the CPU oracle.

The sweep performs the same IEEE f32 operations per update as the single-update kernel, so every comparison
here is bit equality: the current buffer, the previous buffer (recomputed on demand after a sweep), the
on-demand macro field and the info buffer.  Each test also checks through lbm_fused_sweep_count that the
sweep kernel really ran (or, where the host must fall back, that it did not).
"""
import os

import numpy as np
import pytest

import simuverse_b200 as sb
from helpers import assert_bits_equal, tau_default
from simuverse_b200 import wire as W

pytestmark = pytest.mark.gpu


def setting(preset):
    return sb.SettingObj(animation_type=preset)


def node_for(nx, ny, preset, info, flags=0):
    return sb.D2Q9Node((nx * 2, ny * 2), setting(preset), lattice=(nx, ny), lattice_info=info, flags=flags)


def oracle_for(orc, nx, ny, preset, info, threads=4):
    fluid_ty = 1 if preset == W.LID_DRIVEN_CAVITY else 0
    return orc.OracleSim(nx, ny, info, orc.uniform_new(tau_default(), fluid_ty, nx * ny), threads=threads)


def assert_same_state(a, b_dists, b_macro, b_info, what):
    """a: CUDA node; b_*: callables / arrays of the other side."""
    for which in (0, 1):
        assert_bits_equal(a.read_distributions(which), b_dists(which), f"{what} buf{which}")
    assert_bits_equal(a.read_macro(), b_macro(), f"{what} macro")
    assert a.read_lattice_info().tobytes() == b_info().tobytes(), f"{what}: lattice info differs"


E = [(0, 0), (1, 0), (0, -1), (-1, 0), (0, 1), (1, -1), (-1, -1), (-1, 1), (1, 1)]  # fluid/mod.rs:39-49


def live_slots(material, new_solid=None):
    """Slots to compare after solids were painted over live fluid: slot k of a solid cell is only ever read by the
    cell at +e_k; if that one is solid too the slot is dead, and next to a freshly painted solid it is racy in the
    reference itself (boundary.wgsl:28-31 shuffles leftovers between adjacent solids in dispatch order), so dead
    slots are left out (everywhere else both sides hold 0 in them)."""
    ny, nx = material.shape
    solid = (material == W.BOUNDARY) | (material == W.OBSTACLE)
    live = np.ones((9, ny, nx), bool)
    for k in range(1, 9):
        ex, ey = E[k]
        reader_solid = np.roll(solid, (-ey, -ex), axis=(0, 1))  # solid[y + ey, x + ex]
        live[k] = ~(solid & reader_solid)
    return live


def compare_oracle_live(a, sim, live, what):
    assert a.swap_index == sim.swap
    for which in (0, 1):
        got, want = a.read_distributions(which), sim.distributions(which)
        assert_bits_equal(got[live], want[live], f"{what} buf{which} (live slots)")
    assert_bits_equal(a.read_macro(), sim.macro(), f"{what} macro")
    assert a.read_lattice_info().tobytes() == sim.info.tobytes(), f"{what}: lattice info differs"


def compare_nodes(a, b, what):
    assert a.swap_index == b.swap_index
    assert_same_state(a, b.read_distributions, b.read_macro, b.read_lattice_info, what)


def compare_oracle(a, sim, what):
    assert a.swap_index == sim.swap
    assert_same_state(a, sim.distributions, sim.macro, lambda: sim.info, what)


def random_mask(orc, nx, ny, preset, seed, solid=0.08, forces=6):
    """Preset frame + random obstacles + a few permanent force cells (block_iter = -1)."""
    rng = np.random.default_rng(seed)
    info = orc.init_lattice_material(nx, ny, preset)
    g = info.reshape(ny, nx)
    bulk = g["material"] == W.BULK
    g["material"][bulk & (rng.random((ny, nx)) < solid)] = W.OBSTACLE
    ys, xs = np.nonzero(g["material"] == W.BULK)
    for k in rng.choice(len(ys), size=min(forces, len(ys)), replace=False):
        g[ys[k], xs[k]] = (W.EXTERNAL_FORCE, -1, rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1))
    return info


CASES = [
    (600, 375, W.POISEUILLE, 0.0),      # the reference's default lattice: 5 strips, the last one partial
    (128, 128, W.LID_DRIVEN_CAVITY, 0.0),
    (256, 96, W.CUSTOM, 0.10),          # closed box, random obstacles and force cells
    (120, 40, W.POISEUILLE, 0.05),      # exactly one strip: both halo lanes are periodic images
    (124, 33, W.POISEUILLE, 0.05),      # one strip + one group
    (8, 8, W.CUSTOM, 0.0),              # two groups: a strip that wraps onto itself
    (4, 4, W.POISEUILLE, 0.0),          # a single group
    (1024, 24, W.LID_DRIVEN_CAVITY, 0.03),
    (244, 61, W.POISEUILLE, 0.30),      # porous: most cells take the per-cell path
]


@pytest.mark.parametrize("masked", [0, 1])
@pytest.mark.parametrize("nx,ny,preset,solid", CASES)
def test_sweeps_equal_single_updates_and_oracle(orc, nx, ny, preset, solid, masked, monkeypatch):
    """Both instances of the sweep kernel — with (default) and without the inline masked path for pairs next to solids —
    on every case."""
    monkeypatch.setenv("LBM_FUSE_MASKED", str(masked))
    info = random_mask(orc, nx, ny, preset, seed=nx * 1000 + ny, solid=solid) if solid > 0 else \
        orc.init_lattice_material(nx, ny, preset)
    a = node_for(nx, ny, preset, info)
    assert a.sweep_uses_masked_path == bool(masked)
    b = node_for(nx, ny, preset, info, flags=sb.FLAG_NO_FUSE)
    sim = oracle_for(orc, nx, ny, preset, info)
    total = 0
    for n in (2, 1, 7, 40, 33, 16):  # odd counts leave the swap index at 1: sweeps start from either buffer
        a.step_n(n)
        b.step_n(n)
        total += n
        compare_nodes(a, b, f"{nx}x{ny} after {total} updates (vs single updates)")
    sim.step(total)
    compare_oracle(a, sim, f"{nx}x{ny} after {total} updates (vs oracle)")
    assert a.fused_sweep_count >= total // 2 - 3, "the two-update kernel did not run"
    assert b.fused_sweep_count == 0
    a.close()
    b.close()


def test_sweep_kernel_instance_default_and_override(orc, monkeypatch):
    """The instance with the inline masked path is the default; LBM_FUSE_MASKED=0 selects the other one."""
    info = orc.init_lattice_material(256, 128, W.CUSTOM)
    a = node_for(256, 128, W.CUSTOM, info)
    assert a.sweep_uses_masked_path
    a.close()
    monkeypatch.setenv("LBM_FUSE_MASKED", "0")
    a = node_for(256, 128, W.CUSTOM, info)
    assert not a.sweep_uses_masked_path
    a.close()


@pytest.mark.parametrize("rows", [1, 2, 3, 5, 64])
def test_row_block_height_does_not_matter(orc, rows, monkeypatch):
    """Items of H rows recompute rows Y0-1 and Y1 redundantly; any H must give the same lattice."""
    nx, ny = 248, 45
    info = random_mask(orc, nx, ny, W.POISEUILLE, seed=rows, solid=0.06)
    monkeypatch.setenv("LBM_FUSE_ROWS", str(rows))
    a = node_for(nx, ny, W.POISEUILLE, info)
    monkeypatch.delenv("LBM_FUSE_ROWS")
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    a.step_n(50)
    sim.step(50)
    compare_oracle(a, sim, f"H={rows}")
    assert a.fused_sweep_count == 25
    a.close()


def test_frames_are_single_sweeps(orc):
    """lbm_compute_frames without tracer particles: one launch per FluidSimulator::compute frame."""
    nx, ny = 320, 200
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    a = node_for(nx, ny, W.POISEUILLE, info)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    before = a.launch_count
    a.compute_frames(3)
    assert a.launch_count - before == 3 and a.fused_sweep_count == 3
    a.compute_frames(40)  # graph replay of 8 sweeps + single launches
    sim.step(86)
    compare_oracle(a, sim, "43 frames")
    assert a.fused_sweep_count == 43
    a.close()


def test_single_updates_while_a_force_cell_counts_down(orc):
    """collide_stream.wgsl:55-62 mutates the info buffer between updates: no sweeps until the countdown is over."""
    nx, ny = 200, 120
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    a = node_for(nx, ny, W.POISEUILLE, info)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    a.step_n(10)
    sim.step(10)
    assert a.fused_sweep_count == 5
    cells = np.array([(W.EXTERNAL_FORCE, 7, 0.08, -0.05), (W.EXTERNAL_FORCE, 4, -0.02, 0.09)], W.LATTICE_INFO_DTYPE)
    for k, (x, y) in enumerate([(50, 60), (51, 60)]):
        off = (y * nx + x) * 16
        a.write_lattice_info(off, cells[k:k + 1])
        sim.write_lattice_info(off, cells[k:k + 1])
    a.step_n(5)   # still counting
    sim.step(5)
    assert a.fused_sweep_count == 5
    compare_oracle(a, sim, "during the countdown")
    a.step_n(45)  # 3 more single updates (7 + 1 retire), then sweeps
    sim.step(45)
    compare_oracle(a, sim, "after the countdown")
    assert a.fused_sweep_count == 5 + 21
    g = a.read_lattice_info().reshape(ny, nx)
    assert g["material"][60, 50] == W.BULK and g["material"][60, 51] == W.BULK
    a.close()


def test_obstacle_painted_mid_run(orc):
    """An interior obstacle painted over live fluid keeps sweeping; one next to the ghost column leaves values
    that ring cells pull forever (boundary.wgsl:19 never overwrites them), which only single updates reproduce."""
    nx, ny = 240, 160
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    a = node_for(nx, ny, W.POISEUILLE, info)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    a.step_n(20)
    sim.step(20)
    mirror = info.copy()
    off, patch = orc.add_obstacle(nx, ny, mirror, 120, 80)  # d2q9_node.rs:215-245
    a.write_lattice_info(off, patch)
    sim.write_lattice_info(off, patch)
    old_solid = np.isin(info["material"], (W.BOUNDARY, W.OBSTACLE)).reshape(ny, nx)
    live = live_slots(sim.info["material"].reshape(ny, nx), np.isin(sim.info["material"], (W.BOUNDARY, W.OBSTACLE)).reshape(ny, nx) & ~old_solid)
    assert (~live).sum() > 1000
    a.step_n(30)
    sim.step(30)
    compare_oracle_live(a, sim, live, "interior obstacle")
    assert a.fused_sweep_count == 25, "an interior obstacle must not stop the sweeps"
    # solid cells in columns 1 and 2, rows 70..73: the ghost cells (0, 69..74) now pull stale values
    cells = np.zeros(2, W.LATTICE_INFO_DTYPE)
    cells["material"] = W.OBSTACLE
    cells["block_iter"] = -1
    for y in range(70, 74):
        o = (y * nx + 1) * 16
        a.write_lattice_info(o, cells)
        sim.write_lattice_info(o, cells)
    live = live_slots(sim.info["material"].reshape(ny, nx), np.isin(sim.info["material"], (W.BOUNDARY, W.OBSTACLE)).reshape(ny, nx) & ~old_solid)
    a.step_n(30)
    sim.step(30)
    compare_oracle_live(a, sim, live, "obstacle next to the ghost column")
    assert a.fused_sweep_count == 25, "stale values next to the ring: single updates only"
    a.reset()
    sim.reset()
    a.step_n(30)
    sim.step(30)
    compare_oracle(a, sim, "after reset")
    assert a.fused_sweep_count == 40, "lbm_reset zeroes the solids: sweeps again"
    a.close()


def test_restore_then_sweep(orc):
    nx, ny = 160, 96
    info = random_mask(orc, nx, ny, W.CUSTOM, seed=5, solid=0.05)
    a = node_for(nx, ny, W.CUSTOM, info)
    sim = oracle_for(orc, nx, ny, W.CUSTOM, info)
    a.step_n(24)
    sim.step(24)
    saved = [a.read_distributions(0).copy(), a.read_distributions(1).copy()]
    a.step_n(10)
    b = node_for(nx, ny, W.CUSTOM, info)
    b.write_distributions(0, saved[0])
    b.write_distributions(1, saved[1])
    b.step_n(30)
    sim.step(30)
    compare_oracle(b, sim, "restored + 30")
    assert b.fused_sweep_count == 15
    a.close()
    b.close()


def test_stepping_from_the_previous_buffer(orc):
    """lbm_step(swap_index) may name the buffer that is not current; after a sweep it is recomputed first."""
    nx, ny = 128, 64
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    a = node_for(nx, ny, W.POISEUILLE, info)
    b = node_for(nx, ny, W.POISEUILLE, info, flags=sb.FLAG_NO_FUSE)
    for n in (a, b):
        n.step_n(6)
        n.compute_by_pass(1)  # reads buffer 1 (update 5), writes buffer 0
    compare_nodes(a, b, "explicit swap index after sweeps")
    a.close()
    b.close()


def test_large_lattice_sweeps_match_single_updates():
    """BASELINE configs[1] (4096 x 4096 channel + cylinders) at full size: sweeps against single updates."""
    nx = ny = 4096
    a = sb.D2Q9Node((nx, ny), setting(W.POISEUILLE), lattice=(nx, ny), device_preset=W.POISEUILLE)
    b = sb.D2Q9Node((nx, ny), setting(W.POISEUILLE), lattice=(nx, ny), device_preset=W.POISEUILLE, flags=sb.FLAG_NO_FUSE)
    for n in (a, b):
        n.step_n(64)
    assert a.fused_sweep_count == 32
    cur = a.swap_index
    assert_bits_equal(a.read_distributions(cur), b.read_distributions(cur), "4096^2 after 64 updates")
    # k_mass combines per-block partial sums with atomicAdd(double): same bits in, last digits may differ
    assert abs(a.total_mass() - b.total_mass()) <= 1e-12 * b.total_mass()
    a.close()
    b.close()


@pytest.mark.parametrize("n_slabs", [2, 3, 4])
def test_sweeps_on_slabs(orc, n_slabs):
    """Multi-slab lattice: the first / last row block of every slab reads two rows of the neighbour slab over peer
    memory and waits on the same progress flags as the single-update kernel.  Any number of slabs must reproduce the
    oracle; the buffer one update back is recomputed collectively (lbm_refresh_previous)."""
    from simuverse_b200.slabs import SlabGroup

    nx, ny = 248, 96
    info = random_mask(orc, nx, ny, W.POISEUILLE, seed=n_slabs, solid=0.05)
    g = info.reshape(ny, nx)
    for cut in range(1, n_slabs):  # obstacles and a force cell right on the cuts
        y = ny * cut // n_slabs
        g["material"][y - 2:y + 2, 60:75] = W.OBSTACLE
        g[y, 100] = (W.EXTERNAL_FORCE, -1, 0.04, -0.06)
        g[y - 1, 140] = (W.EXTERNAL_FORCE, -1, -0.05, 0.03)
    grp = SlabGroup((nx * 2, ny * 2), setting(W.POISEUILLE), lattice=(nx, ny), n_slabs=n_slabs, lattice_info=info)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    total = 0
    for n in (2, 5, 40, 17):
        grp.step_n(n)
        sim.step(n)
        total += n
        assert grp.swap_index == sim.swap
        for which in (0, 1):
            assert_bits_equal(grp.read_distributions(which), sim.distributions(which), f"{n_slabs} slabs, {total} updates, buf{which}")
        assert_bits_equal(grp.read_macro(), sim.macro(), f"{n_slabs} slabs, {total} updates, macro")
    assert all(n.fused_sweep_count == 31 for n in grp.nodes), [n.fused_sweep_count for n in grp.nodes]
    # an interior mask write after reset: every slab keeps sweeping (the verdict comes from the bytes every rank sees)
    cells = np.zeros(3, W.LATTICE_INFO_DTYPE)
    cells["material"] = W.OBSTACLE
    cells["block_iter"] = -1
    off = (50 * nx + 120) * 16
    grp.write_lattice_info(off, cells)
    sim.write_lattice_info(off, cells)
    grp.step_n(10)
    sim.step(10)
    live = live_slots(sim.info["material"].reshape(ny, nx))
    for which in (0, 1):
        got, want = grp.read_distributions(which), sim.distributions(which)
        assert_bits_equal(got[live], want[live], f"after a mask write, buf{which}")
    assert all(n.fused_sweep_count == 36 for n in grp.nodes), [n.fused_sweep_count for n in grp.nodes]
    grp.close()


@pytest.mark.parametrize("preset,solid", [(W.POISEUILLE, 0.30), (W.CUSTOM, 0.08)])
def test_divergent_fallback_paths_on_a_wide_lattice(orc, preset, solid):
    """4096 columns (18 CTA columns, a partial last strip), solids everywhere and hundreds of force cells whose numerators
    hit the exact-division fallback (denormal components) or the signed-zero rule: every warp mixes the vector path,
    the masked path, the out-of-line per-cell path (outer ring, first / last column) and the division fallback.  This
    is the configuration on which a call from a divergent branch once made a lane run the row loop on its own."""
    nx, ny = 4096, 96
    rng = np.random.default_rng(17)
    info = random_mask(orc, nx, ny, preset, seed=23, solid=solid, forces=0)
    g = info.reshape(ny, nx)
    ys, xs = np.nonzero(g["material"] == W.BULK)
    for k in rng.choice(len(ys), size=400, replace=False):
        vx = rng.choice([1e-39, -1e-40, -0.0, 0.02, -0.03])
        vy = rng.choice([-1e-39, 3e-41, -0.0, 0.0, 0.04])
        g[ys[k], xs[k]] = (W.EXTERNAL_FORCE, -1, vx, vy)
    a = node_for(nx, ny, preset, info, flags=sb.FLAG_MACRO_EVERY_STEP)
    b = node_for(nx, ny, preset, info, flags=sb.FLAG_MACRO_EVERY_STEP | sb.FLAG_NO_FUSE)
    sim = oracle_for(orc, nx, ny, preset, info, threads=orc.lib().orc_get_max_threads())
    for n in (6, 31):
        a.step_n(n)
        b.step_n(n)
        sim.step(n)
        cur = a.swap_index
        assert_bits_equal(a.read_distributions(cur), b.read_distributions(cur), f"sweeps vs single updates after +{n}")
        assert_bits_equal(a.read_distributions(cur), sim.distributions(cur), f"sweeps vs oracle after +{n}")
        np.testing.assert_array_equal(a.read_macro_tex().view(np.uint16).reshape(-1), sim.macro_f16)
    assert a.fused_sweep_count == 18 and b.fused_sweep_count == 0
    a.close()
    b.close()


@pytest.mark.parametrize("nx,n_slabs,rows", [(520, 1, 0), (520, 3, 0), (760, 2, 0), (520, 4, 4), (1000, 1, 3)])
def test_packed_last_strip_column(orc, monkeypatch, nx, n_slabs, rows):
    """The last CTA column of a sweep holds one or two strips at these widths (520 -> 9 strips, 760 -> 13 -> one; 1000 -> 17 ->
    one): its CTAs take four / two row blocks each (FuseGeom::pack_a).  LBM_FUSE_PACK=2 forces the packing on lattices
    that fit in one wave, where it is normally off; on slabs the first packed CTA holds the edge row blocks: it waits
    for the neighbours as a whole and every warp signals for its own row block."""
    from simuverse_b200.slabs import SlabGroup

    monkeypatch.setenv("LBM_FUSE_PACK", "2")
    if rows:
        monkeypatch.setenv("LBM_FUSE_ROWS", str(rows))
    ny = 96
    info = random_mask(orc, nx, ny, W.POISEUILLE, seed=nx + n_slabs, solid=0.04)
    g = info.reshape(ny, nx)
    for cut in range(1, n_slabs):  # solids and force cells on the cuts, also inside the last strip column
        y = ny * cut // n_slabs
        g["material"][y - 2:y + 2, nx - 40:nx - 25] = W.OBSTACLE
        g[y, nx - 12] = (W.EXTERNAL_FORCE, -1, 0.04, -0.06)
        g[y - 1, nx - 50] = (W.EXTERNAL_FORCE, -1, -0.05, 0.03)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    if n_slabs == 1:
        a = node_for(nx, ny, W.POISEUILLE, info)
        nodes = [a]
    else:
        a = SlabGroup((nx * 2, ny * 2), setting(W.POISEUILLE), lattice=(nx, ny), n_slabs=n_slabs, lattice_info=info)
        nodes = a.nodes
    total = 0
    for n in (2, 7, 40):
        a.step_n(n)
        sim.step(n)
        total += n
        for which in (0, 1):
            assert_bits_equal(a.read_distributions(which), sim.distributions(which), f"{nx} wide, {n_slabs} slab(s), {total} updates, buf{which}")
        assert_bits_equal(a.read_macro(), sim.macro(), f"{total} updates, macro")
    assert all(n.fused_sweep_count == 24 for n in nodes), [n.fused_sweep_count for n in nodes]
    a.close()
