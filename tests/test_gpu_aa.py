"""AA-pattern in-place variant (LBM_FLAG_AA): one copy of the distributions, two alternating layouts.
After any number of updates the canonicalised state must equal the reference's current buffer bit for
bit — the same bar as for the A/B ping-pong kernels."""
import numpy as np
import pytest

import simuverse_b200 as sb
from helpers import assert_bits_equal, tau_default
from simuverse_b200 import wire as W

pytestmark = pytest.mark.gpu


def setting(preset):
    return sb.SettingObj(animation_type=preset)


def pair(orc, nx, ny, preset, info=None, flags=0):
    if info is None:
        info = orc.init_lattice_material(nx, ny, preset)
    fluid_ty = 1 if preset == W.LID_DRIVEN_CAVITY else 0
    node = sb.D2Q9Node((nx * 2, ny * 2), setting(preset), lattice=(nx, ny), lattice_info=info, flags=sb.FLAG_AA | flags)
    sim = orc.OracleSim(nx, ny, info, orc.uniform_new(tau_default(), fluid_ty, nx * ny), threads=4)
    return node, sim


def check(node, sim, what):
    assert node.swap_index == sim.swap
    assert_bits_equal(node.read_distributions(node.swap_index), sim.distributions(sim.swap), what)
    assert node.read_lattice_info().tobytes() == sim.info.tobytes()
    assert abs(node.total_mass() - sim.total_mass()) <= 1e-12 * sim.total_mass()


@pytest.mark.parametrize("nx,ny,preset", [
    (600, 375, W.POISEUILLE), (128, 128, W.LID_DRIVEN_CAVITY), (131, 77, W.POISEUILLE), (37, 19, W.CUSTOM),
    (513, 40, W.LID_DRIVEN_CAVITY), (1030, 9, W.CUSTOM), (3, 3, W.CUSTOM),
])
def test_aa_matches_oracle_at_odd_and_even_steps(orc, nx, ny, preset):
    node, sim = pair(orc, nx, ny, preset)
    check(node, sim, "after init")
    for k in (1, 1, 1, 8, 35, 54):  # 1, 2, 3, 11, 46, 100 updates: both layouts get read back
        node.step_n(k)
        sim.step(k)
        check(node, sim, f"{nx}x{ny} after +{k}")
    node.close()


def test_aa_all_fluid_periodic_and_force_cells(orc):
    nx, ny = 260, 37
    info = np.zeros(nx * ny, W.LATTICE_INFO_DTYPE)
    info["material"], info["block_iter"] = W.BULK, -1
    g = info.reshape(ny, nx)
    g[0, 0] = (W.EXTERNAL_FORCE, -1, 0.06, 0.05)
    g["material"][15:19, 100:104] = W.OBSTACLE
    g["material"][0, 200:204] = W.OBSTACLE          # solids on the ring: ring neighbours read their stale slots
    g["material"][10:13, nx - 1] = W.OBSTACLE       # and across the periodic wrap
    node, sim = pair(orc, nx, ny, W.CUSTOM, info=info)
    cell = np.zeros(1, W.LATTICE_INFO_DTYPE)
    cell[0] = (W.EXTERNAL_FORCE, 7, -0.05, 0.03)     # armed after init: counts down and flips on the device
    node.write_lattice_info((20 * nx + 50) * 16, cell)
    sim.write_lattice_info((20 * nx + 50) * 16, cell)
    for k in (1, 5, 1, 1, 1, 40, 41):
        node.step_n(k)
        sim.step(k)
        check(node, sim, f"periodic +{k}")
    node.close()


def test_aa_porous_and_macro_texture(orc):
    nx, ny = 384, 200
    info = orc.init_porous_material(nx, ny)
    node, sim = pair(orc, nx, ny, W.POISEUILLE, info=info, flags=sb.FLAG_MACRO_EVERY_STEP)
    for k in (1, 30, 30):
        node.step_n(k)
        sim.step(k)
        check(node, sim, f"porous +{k}")
        np.testing.assert_array_equal(node.read_macro_tex().view(np.uint16).reshape(-1), sim.macro_f16)
    node.close()


def test_aa_frames_with_particles_match_ab(orc):
    canvas = (400, 260)
    fa = sb.FluidSimulator(canvas, setting(W.POISEUILLE), particles=True, flags=sb.FLAG_AA)
    fb = sb.FluidSimulator(canvas, setting(W.POISEUILLE), particles=True)
    fa.compute(30)
    fb.compute(30)
    na, nb = fa.fluid_compute_node, fb.fluid_compute_node
    n = fa.particles_num[0] * fa.particles_num[1]
    assert na.read_particles(n).tobytes() == nb.read_particles(n).tobytes()
    assert_bits_equal(na.read_distributions(na.swap_index), nb.read_distributions(nb.swap_index), "AA vs A/B frames")
    na.close()
    nb.close()


def test_aa_restrictions_are_reported(orc):
    nx, ny = 96, 64
    node, sim = pair(orc, nx, ny, W.POISEUILLE)
    node.step_n(1)
    with pytest.raises(sb.LbmError) as e:
        node.read_distributions(0)          # the previous buffer does not exist
    assert e.value.status == 5
    with pytest.raises(sb.LbmError):
        node.read_macro()                    # the step overwrote its inputs
    with pytest.raises(sb.LbmError):
        node.compute_by_pass(0)              # state is in the shifted layout: next step is swap_index 1
    with pytest.raises(sb.LbmError):
        node.write_distributions(0, np.zeros((9, ny, nx), np.float32))
    node.compute_by_pass(1)
    sim.step(2)
    check(node, sim, "after explicit swap")
    # checkpoint / restore in the natural layout
    snap = node.read_distributions(0)
    node.step_n(10)
    want = node.read_distributions(node.swap_index)
    node2 = sb.D2Q9Node((nx * 2, ny * 2), setting(W.POISEUILLE), lattice=(nx, ny), lattice_info=node.read_lattice_info(),
                        flags=sb.FLAG_AA)
    node2.write_distributions(0, snap)
    node2.step_n(10)
    assert_bits_equal(node2.read_distributions(node2.swap_index), want, "AA restore")
    with pytest.raises(sb.LbmError):
        sb.D2Q9Node((nx * 2, ny * 2), setting(W.POISEUILLE), lattice=(nx, ny), flags=sb.FLAG_AA, rank=0, world=2)
    node.close()
    node2.close()
