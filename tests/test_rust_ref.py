"""The reference's Rust HOST helpers on the LBM path, pinned on their EXECUTED source text.

tests/golden/rust_host_helpers.npz was produced by running the function bodies of fluid/lattice.rs, fluid/mod.rs,
fluid/d2q9_node.rs and fluid/fluid_simulator.rs through tests/rust_ref (a Rust-subset transpiler; the image has no Rust
toolchain).  Here the oracle (oracle/lbm_oracle.c) and the product's host side (csrc/host_logic.cpp through the C ABI,
and the Python mirrors of D2Q9Node / FluidSimulator driven without a device) must reproduce every byte of it; when the
reference tree is present the shaders' host counterparts are re-executed live on a subset."""
import ctypes as C
import os
import types

import numpy as np
import pytest

import simuverse_b200 as sb
from helpers import GOLDEN_DIR
from simuverse_b200 import wire as W
from simuverse_b200.wire import ptr

G = np.load(os.path.join(GOLDEN_DIR, "rust_host_helpers.npz"))
NX, NY, LPS = (int(v) for v in G["lattice"])


def test_masks_match_executed_rust(orc):
    """init_lattice_material (fluid/lattice.rs:26-98), every preset"""
    for k, (nx, ny, ty) in enumerate(G["masks"]):
        want = G[f"mask_{k}"].tobytes()
        assert orc.init_lattice_material(int(nx), int(ny), int(ty)).tobytes() == want, f"oracle mask {nx}x{ny} ty {ty}"
        assert sb.init_lattice_material(int(nx), int(ny), int(ty)).tobytes() == want, f"product mask {nx}x{ny} ty {ty}"


def test_mask_of_baseline_config_2_matches_executed_rust(orc):
    """4096 x 4096 Poiseuille mask (BASELINE configs[1]): CRC32 of the i32-LE material array, histogram and the inlet's
    vx as produced by executing init_lattice_material's text over all 16.7 M cells (make_rust_golden.py --crc4096)."""
    import zlib

    for gen in (orc.init_lattice_material, sb.init_lattice_material):
        m = gen(4096, 4096, W.POISEUILLE)
        assert zlib.crc32(m["material"].astype("<i4").tobytes()) == int(G["mask4096_crc"]) == 0x63A17811
        np.testing.assert_array_equal(np.bincount(m["material"], minlength=8), G["mask4096_hist"])
        assert float(m["vx"].astype(np.float64).sum()) == float(G["mask4096_vx_sum"])
        assert (m["block_iter"] == -1).all() and (m["vy"] == 0).all()


def test_uniforms_match_executed_rust(orc):
    """LbmUniform::new (fluid/mod.rs:32-55) and update_uniforms (fluid_simulator.rs:175-193)"""
    want = G["uniform_bytes"].tobytes()
    for k, (tau, ty, n) in enumerate(G["uniform_args"]):
        w = want[304 * k:304 * (k + 1)]
        assert bytes(orc.uniform_new(float(np.float32(tau)), int(ty), int(n))) == w
        assert bytes(sb.lbm_uniform_new(float(np.float32(tau)), int(ty), int(n))) == w
    want = G["update_uniform_bytes"].tobytes()
    for k, (v, ty) in enumerate(G["viscosities"]):
        w = want[304 * k:304 * (k + 1)]
        fluid_ty = 1 if int(ty) == W.LID_DRIVEN_CAVITY else 0
        assert bytes(orc.uniform_new(orc.tau_from_viscosity(v), fluid_ty, NX * NY)) == w
        # the product's FluidSimulator.update_uniforms, run on a stand-in that records the uniform it would upload
        rec = []
        stub = types.SimpleNamespace(lattice=(NX, NY),
                                     fluid_compute_node=types.SimpleNamespace(write_uniform=lambda u: rec.append(bytes(u))))
        setting = types.SimpleNamespace(fluid_viscosity=float(np.float32(v)), animation_type=int(ty))
        sb.FluidSimulator.update_uniforms(stub, setting)
        assert rec == [w]


class _NodeStub:
    """The host-only part of simuverse_b200.D2Q9Node: its mask painters run against a write recorder, no device."""

    def __init__(self, info):
        self.lattice = (NX, NY)
        self.lattice_pixel_size = LPS
        self.lattice_info_data = info.copy()
        self.writes = []

    def _mirror(self):
        return self.lattice_info_data

    def write_lattice_info(self, byte_offset, cells):
        self.writes.append((int(byte_offset), np.ascontiguousarray(cells, W.LATTICE_INFO_DTYPE).tobytes()))

    add_obstacle = sb.D2Q9Node.add_obstacle
    add_external_force = sb.D2Q9Node.add_external_force


def _sim_stub(info):
    node = _NodeStub(info)
    return types.SimpleNamespace(lattice=(NX, NY), lattice_pixel_size=LPS, pre_pos=(0.0, 0.0), fluid_compute_node=node), node


def test_clicks_match_executed_rust(orc):
    """on_click -> add_obstacle (fluid_simulator.rs:137-152, d2q9_node.rs:215-245)"""
    sim, node = _sim_stub(G["mask_0"])
    mirror = G["mask_0"].copy()
    k = 0
    for pos, wrote in zip(G["clicks"], G["click_wrote"]):
        n0 = len(node.writes)
        assert sb.FluidSimulator.on_click(sim, (float(pos[0]), float(pos[1]))) == bool(wrote)
        assert (len(node.writes) > n0) == bool(wrote)
        cell = orc.on_click_guard(NX, NY, LPS, float(pos[0]), float(pos[1]))
        assert (cell is not None) == bool(wrote)
        if wrote:
            off, patch = orc.add_obstacle(NX, NY, mirror, *cell)
            assert off == int(G["click_offsets"][k]) and patch.tobytes() == G["click_patches"][k].tobytes()
            assert node.writes[-1] == (int(G["click_offsets"][k]), G["click_patches"][k].tobytes())
            k += 1
    assert k == len(G["click_offsets"])
    assert mirror.tobytes() == G["mirror_after_clicks"].tobytes()
    assert node.lattice_info_data.tobytes() == G["mirror_after_clicks"].tobytes()


def test_drag_matches_executed_rust(orc):
    """touch_begin / touch_move -> add_external_force (fluid_simulator.rs:154-173, d2q9_node.rs:263-300): the pre_pos
    bookkeeping and every 16-byte write, incl. the atan2 / cos / sin force components"""
    sim, node = _sim_stub(G["mask_0"])
    sb.FluidSimulator.touch_begin(sim)
    pre = (0.0, 0.0)
    at = 0
    for pos, count, want_pre in zip(G["drag"], G["drag_write_counts"], G["drag_pre_pos"]):
        n0 = len(node.writes)
        sb.FluidSimulator.touch_move(sim, (float(pos[0]), float(pos[1])))
        assert len(node.writes) - n0 == int(count)
        assert tuple(np.float32(sim.pre_pos)) == tuple(want_pre)
        if count:  # the oracle's add_external_force with the golden's own pre_pos
            offs, cells = orc.add_external_force(NX, NY, LPS, (float(pos[0]), float(pos[1])), pre)
            np.testing.assert_array_equal(offs, G["drag_offsets"][at:at + count])
            assert cells.tobytes() == G["drag_cells"][at:at + count].tobytes()
        at += int(count)
        pre = (float(want_pre[0]), float(want_pre[1]))
    assert [w[0] for w in node.writes] == [int(o) for o in G["drag_offsets"]]
    assert b"".join(w[1] for w in node.writes) == G["drag_cells"].tobytes()
    assert node.lattice_info_data.tobytes() == G["mask_0"].tobytes()  # force writes never touch the CPU mirror


def test_particle_grid_matches_executed_rust(orc):
    """get_particles_data (lib.rs:247-264): tracer grid extent; the reference dispatches ceil(extent / 16) workgroups"""
    for (w, h, count), extent, groups in zip(G["grid_args"], G["grid_extent"], G["grid_workgroups"]):
        assert orc.particle_grid(int(w), int(h), int(count)) == tuple(extent)
        assert sb.particle_grid((int(w), int(h)), int(count)) == tuple(extent)
        assert (-(-extent[0] // 16), -(-extent[1] // 16), 1) == tuple(groups)


def test_field_uniform_matches_executed_rust(orc):
    """FieldUniform as D2Q9Node::new builds it (d2q9_node.rs:38-76, util/matrix_helper.rs:19-41): all 48 bytes, including
    proj_ratio / ndc_pixel, which no LBM shader reads — executing the reference's text is what showed that the oracle and
    the product used to leave them at zero."""
    for k, (cw, ch, lps) in enumerate(G["field_args"]):
        want = G["field_bytes"].tobytes()[48 * k:48 * (k + 1)]
        nx, ny = int(cw) // int(lps), int(ch) // int(lps)
        assert bytes(orc.field_uniform_new(nx, ny, int(lps), int(cw), int(ch))) == want
        f = W.FieldUniform()
        sb.lib.lbm_field_uniform_new(nx, ny, int(lps), int(cw), int(ch), C.byref(f))
        assert bytes(f) == want


def test_live_execution_of_the_rust_source_matches_the_golden():
    """Re-executes the reference's Rust text when the tree is present (build container): small masks, one click, the drag."""
    from rust_ref import harness as H

    if not H.available():
        pytest.skip("reference tree not present (GPU box)")
    for k, (nx, ny, ty) in enumerate(G["masks"]):
        if nx * ny <= 131 * 77:
            assert H.init_lattice_material(int(nx), int(ny), int(ty)).tobytes() == G[f"mask_{k}"].tobytes()
    for k, (tau, ty, n) in enumerate(G["uniform_args"]):
        assert H.uniform_new(tau, int(ty), int(n)) == G["uniform_bytes"].tobytes()[304 * k:304 * (k + 1)]
    for (w, h, count), extent in zip(G["grid_args"], G["grid_extent"]):
        assert H.particle_grid(int(w), int(h), int(count))[0] == tuple(extent)
    for k, (cw, ch, lps) in enumerate(G["field_args"]):
        assert H.field_uniform(int(cw), int(ch), int(lps)) == G["field_bytes"].tobytes()[48 * k:48 * (k + 1)]
    sim = H.Simulator(NX, NY, LPS, W.POISEUILLE, G["mask_0"])
    first = int(np.nonzero(G["click_wrote"])[0][0])
    sim.on_click(*G["clicks"][first])
    assert sim.writes[-1][1:] == (int(G["click_offsets"][0]), G["click_patches"][0].tobytes())
    del sim.writes[:]
    sim.touch_begin()
    for pos in G["drag"]:
        sim.touch_move(*pos)
    assert [w[1] for w in sim.writes] == [int(o) for o in G["drag_offsets"]]
    assert b"".join(w[2] for w in sim.writes) == G["drag_cells"].tobytes()


def test_the_repos_rust_shim_executed_leaves_the_reference_writes():
    """rust/simuverse-cuda-lbm (CudaD2Q9Node / CudaFluidSimulator) is shipped as source that this image cannot compile.
    Its host logic — written independently of the reference's wording — is executed by the same Rust-subset interpreter,
    with the FFI calls replaced by recorders: clicks, the drag and the uniform updates of the golden must leave exactly
    the writes the reference's own Rust text leaves."""
    from rust_ref import harness as H

    if not H.available():
        pytest.skip("reference tree not present (the shim imports the reference's own helper items)")
    sim = H.ShimSimulator(NX, NY, LPS, W.POISEUILLE, G["mask_0"])
    for v, ty in G["viscosities"]:
        sim.update_uniforms(v, int(ty))
    assert b"".join(w[2] for w in sim.writes) == G["update_uniform_bytes"].tobytes()
    del sim.writes[:]
    k = 0
    for pos, wrote in zip(G["clicks"], G["click_wrote"]):
        n0 = len(sim.writes)
        sim.on_click(*pos)
        assert (len(sim.writes) > n0) == bool(wrote)
        if wrote:
            assert sim.writes[-1] == ("info_buf", int(G["click_offsets"][k]), G["click_patches"][k].tobytes())
            k += 1
    assert H.info_to_array(sim.fluid_compute_node.lattice_info_data).tobytes() == G["mirror_after_clicks"].tobytes()
    del sim.writes[:]
    sim.touch_begin()
    for pos, count, want_pre in zip(G["drag"], G["drag_write_counts"], G["drag_pre_pos"]):
        n0 = len(sim.writes)
        sim.touch_move(*pos)
        assert len(sim.writes) - n0 == int(count)
        assert (sim.pre_pos.x, sim.pre_pos.y) == tuple(want_pre)
    assert [w[1] for w in sim.writes] == [int(o) for o in G["drag_offsets"]]
    assert b"".join(w[2] for w in sim.writes) == G["drag_cells"].tobytes()
