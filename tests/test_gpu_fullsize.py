"""Independent parity at BASELINE.json's full sizes (round 2): every comparison here is against the CPU oracle,
none against another kernel of this repo.

* 4096 x 4096 (configs[1]): the whole lattice, 100 updates, bit for bit (north_star's "after 100 steps").
* 16384 x 16384 (configs[2]) on 8 slabs and 8192 x 8192 porous + 1M tracers (configs[4]): the oracle cannot hold
  those lattices for long, so it runs BANDS of them.  A window of rows [a, b) after k updates depends only on rows
  [a-k, b+k) of the initial state (one row per update in each direction); the oracle is run on the window plus
  k+2 rows of margin of the same mask — taken from the oracle's own generator, against which the device-generated
  mask is compared byte for byte first — and the window rows must agree bitwise.  Windows: the two walls, every slab
  cut (the rows that travel over peer memory), full width (both ghost columns, the inlet and the outlet).  With
  nx * ny = 2^28 sites and 9 planes these runs are also the check that cell indexing is 64-bit
  (d2q9_fn.wgsl:13-17 is i32 in the reference).
"""
import zlib

import numpy as np
import pytest

import simuverse_b200 as sb
from helpers import assert_bits_equal, assert_close_rel, tau_default
from simuverse_b200 import wire as W

pytestmark = pytest.mark.gpu


def setting():
    return sb.SettingObj(animation_type=W.POISEUILLE)


def band_oracle(orc, info2d, a, b, k):
    """Oracle over rows [a - k - 2, b + k + 2) (clipped to the lattice) of mask info2d, k updates from init.
    Returns (sim, first row of the band)."""
    ny, nx = info2d.shape
    lo, hi = max(0, a - k - 2), min(ny, b + k + 2)
    band = np.ascontiguousarray(info2d[lo:hi]).reshape(-1)
    sim = orc.OracleSim(nx, hi - lo, band, orc.uniform_new(tau_default(), 0, (nx * (hi - lo)) & 0x7FFFFFFF),
                        threads=orc.lib().orc_get_max_threads())
    sim.step(k)
    return sim, lo


def test_4096_100_updates_against_the_oracle(orc):
    nx = ny = 4096
    node = sb.D2Q9Node((nx * 2, ny * 2), setting(), lattice=(nx, ny), device_preset=W.POISEUILLE,
                       flags=sb.FLAG_MACRO_EVERY_STEP)
    info = node.read_lattice_info()
    assert zlib.crc32(info["material"].astype("<i4").tobytes()) == 0x63A17811
    want_info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    assert info.tobytes() == want_info.tobytes(), "device-generated mask differs from init_lattice_material"
    sim = orc.OracleSim(nx, ny, want_info, orc.uniform_new(tau_default(), 0, nx * ny), threads=orc.lib().orc_get_max_threads())
    node.step_n(100)
    sim.step(100)
    assert node.fused_sweep_count == 50
    for which in (0, 1):
        got, want = node.read_distributions(which), sim.distributions(which)
        assert_bits_equal(got, want, f"4096^2, 100 updates, buf{which}")
        assert_close_rel(got, want, rel=1e-5, what="north_star tolerance (implied by bit equality)")
    np.testing.assert_array_equal(node.read_macro_tex().view(np.uint16).reshape(-1), sim.macro_f16)
    assert_bits_equal(node.read_macro(), sim.macro(), "(u, rho) of update 100")
    assert abs(node.total_mass() - sim.total_mass()) / sim.total_mass() < 1e-12
    cur = node.read_distributions(node.swap_index)
    mx = [0.6] + [0.2222] * 4 + [0.1111] * 4
    for i in range(9):
        assert cur[i].min() >= 0.0 and cur[i].max() <= np.float32(mx[i])
    node.close()


def test_16384_on_8_slabs_band_checks_against_the_oracle(orc):
    from simuverse_b200.slabs import SlabGroup

    nx = ny = 16384
    n_slabs, k, half = 8, 12, 32
    grp = SlabGroup((nx * 2, ny * 2), setting(), lattice=(nx, ny), n_slabs=n_slabs, device_preset=W.POISEUILLE)
    want_info = orc.init_lattice_material(nx, ny, W.POISEUILLE).reshape(ny, nx)
    h = ny // n_slabs
    for r, n in enumerate(grp.nodes):
        assert n.read_lattice_info().tobytes() == want_info[r * h:(r + 1) * h].tobytes(), f"mask of slab {r}"
    grp.step_n(k)
    grp.sync()
    assert all(n.fused_sweep_count == k // 2 for n in grp.nodes)
    swap = grp.swap_index
    # first and last 2*half rows of every slab (current buffer), then the slabs can go
    edges = {}
    for r, n in enumerate(grp.nodes):
        d = n.read_distributions(swap)
        edges[r] = (d[:, :2 * half].copy(), d[:, -2 * half:].copy())
        del d
    mass = grp.total_mass()
    grp.close()
    windows = [(0, 2 * half), (ny - 2 * half, ny)] + [(c * h - half, c * h + half) for c in range(1, n_slabs)]
    for a, b in windows:
        sim, lo = band_oracle(orc, want_info, a, b, k)
        want = sim.distributions(sim.swap)[:, a - lo:b - lo]
        if a == 0:
            got = edges[0][0]
        elif b == ny:
            got = edges[n_slabs - 1][1]
        else:
            c = b // h  # window straddles the cut between slabs c-1 and c
            got = np.concatenate([edges[c - 1][1][:, half:], edges[c][0][:, :half]], axis=1)
        assert sim.swap == swap
        assert_bits_equal(got, want, f"16384^2 on {n_slabs} slabs, {k} updates, rows [{a}, {b})")
    assert 0.99 * nx * ny < mass < 1.01 * nx * ny


def test_8192_porous_with_1m_tracers_band_checks_against_the_oracle(orc):
    """BASELINE configs[4] as bench.py --config 5 runs it: FluidSimulator frames (two-update sweeps through the masked
    path + two particle passes of 1000 x 1000 tracers)."""
    nx = ny = 8192
    frames = 6
    k = 2 * frames
    s = sb.SettingObj(animation_type=W.POISEUILLE, particles_count=1000000)
    fs = sb.FluidSimulator((nx * 2, ny * 2), s, particles=True, lattice=(nx, ny), device_preset=sb.PRESET_POROUS)
    node = fs.fluid_compute_node
    want_info = orc.init_porous_material(nx, ny).reshape(ny, nx)
    assert node.read_lattice_info().tobytes() == want_info.tobytes(), "device-generated porous mask"
    solid = (want_info["material"] == W.OBSTACLE).mean()
    assert 0.29 < solid < 0.31
    fs.compute(frames)
    assert node.fused_sweep_count == frames
    cur = node.read_distributions(node.swap_index)
    tex = node.read_macro_tex().view(np.uint16)
    parts = node.read_particles(fs.particles_num[0] * fs.particles_num[1])
    assert np.isfinite(parts["pos"]).all() and (parts["pos"] != parts["pos_initial"]).any()
    node.close()
    for a, b in [(0, 64), (ny // 2 - 32, ny // 2 + 32), (5000, 5064), (ny - 64, ny)]:
        sim, lo = band_oracle(orc, want_info, a, b, k)
        assert_bits_equal(cur[:, a:b], sim.distributions(sim.swap)[:, a - lo:b - lo], f"8192^2 porous rows [{a}, {b})")
        np.testing.assert_array_equal(tex[a:b].reshape(-1), sim.macro_f16.reshape(-1, nx * 4)[a - lo:b - lo].reshape(-1),
                                      err_msg=f"macro texture rows [{a}, {b})")
