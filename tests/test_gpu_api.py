"""GPU tests of the C-ABI surface beyond the straight parity runs: launch regimes of the step
(inline-mixed / mixed-warp list / mixed-everywhere), CUDA-graph replay, checkpoint/restore, partial
LatticeInfo writes, call-order errors, ragged and minimal lattices, full-size properties."""
import ctypes as C

import numpy as np
import pytest

import simuverse_b200 as sb
from helpers import assert_bits_equal, tau_default
from simuverse_b200 import _capi
from simuverse_b200 import wire as W
from simuverse_b200._capi import lib
from simuverse_b200.wire import ptr

pytestmark = pytest.mark.gpu


def setting(preset):
    return sb.SettingObj(animation_type=preset)


def oracle_for(orc, nx, ny, info, fluid_ty=0, threads=4):
    return orc.OracleSim(nx, ny, info, orc.uniform_new(tau_default(), fluid_ty, nx * ny), threads=threads)


def striped_mask(orc, nx, ny, period, width=3):
    """Custom frame + a small solid block in every `period`-th 128-cell span of every 3rd row band:
    controls the fraction of mixed warps."""
    info = orc.init_lattice_material(nx, ny, W.CUSTOM)
    g = info.reshape(ny, nx)
    for w in range(0, nx // 128):
        if w % period == 0:
            x = w * 128 + 40
            g["material"][2:ny - 2:1, x:x + width] = W.OBSTACLE
    g[ny // 2, 10:20] = (W.EXTERNAL_FORCE, -1, 0.04, 0.02)
    return info


@pytest.mark.parametrize("period,regime", [(16, "rare"), (4, "list"), (1, "everywhere")])
def test_launch_regimes_match_oracle(orc, period, regime):
    """<= 1/8 mixed warps: one inline launch; <= 1/2: pure kernel + list kernel; else mixed kernel over
    everything.  All three must leave identical buffers."""
    nx, ny = 4096, 384  # 32 warps per row: ~3/32, ~9/32 and 32/32 of them mixed; > 8192 interior warps, so the
    # small-lattice rule (always inline) does not apply
    info = striped_mask(orc, nx, ny, period)
    node = sb.D2Q9Node((nx * 2, ny * 2), setting(W.CUSTOM), lattice=(nx, ny), lattice_info=info, flags=sb.FLAG_NO_FUSE)
    node.step_n(1)  # the first single update after a mask change also rebuilds the mixed-warp list (k_scan_mixed)
    before = node.launch_count
    node.step_n(3)
    per_step = (node.launch_count - before) / 3
    assert per_step == (1 if regime == "rare" else 2), f"{regime}: {per_step} launches per update"
    node.step_n(56)
    sim = oracle_for(orc, nx, ny, info)
    sim.step(60)
    for which in (0, 1):
        assert_bits_equal(node.read_distributions(which), sim.distributions(which), f"{regime} buf{which}")
    assert_bits_equal(node.read_macro(), sim.macro(), f"{regime} macro")
    node.close()


@pytest.mark.parametrize("nx,ny", [(3, 3), (4, 5), (5, 4), (7, 9), (8, 8), (127, 6), (128, 6), (129, 6), (130, 5),
                                    (255, 4), (512, 3), (516, 7), (1030, 5)])
def test_ragged_and_minimal_lattices(orc, nx, ny):
    info = orc.init_lattice_material(nx, ny, W.CUSTOM)
    g = info.reshape(ny, nx)
    if nx > 6 and ny > 4:
        g[ny // 2, nx // 2] = (W.EXTERNAL_FORCE, -1, 0.05, -0.03)
    for flags in (0, sb.FLAG_KERNEL_GENERIC, sb.FLAG_MACRO_EVERY_STEP):
        node = sb.D2Q9Node((nx * 2, ny * 2), setting(W.CUSTOM), lattice=(nx, ny), lattice_info=info, flags=flags)
        sim = oracle_for(orc, nx, ny, info, threads=1)
        node.step_n(0)  # empty request is a no-op
        node.step_n(41)
        sim.step(41)
        for which in (0, 1):
            assert_bits_equal(node.read_distributions(which), sim.distributions(which), f"{nx}x{ny} flags={flags}")
        if flags & sb.FLAG_MACRO_EVERY_STEP:
            np.testing.assert_array_equal(node.read_macro_tex().view(np.uint16).reshape(-1), sim.macro_f16)
        node.close()


def test_all_fluid_periodic_lattice(orc):
    """No solids at all: every pull wraps periodically in x and y (layout_and_fn.wgsl:38-51)."""
    nx, ny = 260, 37
    info = np.zeros(nx * ny, W.LATTICE_INFO_DTYPE)
    info["material"], info["block_iter"] = W.BULK, -1
    g = info.reshape(ny, nx)
    g[0, 0] = (W.EXTERNAL_FORCE, -1, 0.06, 0.05)
    g[ny - 1, nx - 1] = (W.EXTERNAL_FORCE, 25, -0.05, 0.02)  # armed before init: disarmed by init.wgsl:51-59
    node = sb.D2Q9Node((nx * 2, ny * 2), setting(W.CUSTOM), lattice=(nx, ny), lattice_info=info)
    sim = oracle_for(orc, nx, ny, info)
    node.step_n(90)
    sim.step(90)
    assert_bits_equal(node.read_distributions(node.swap_index), sim.distributions(sim.swap), "periodic")
    assert node.read_lattice_info().tobytes() == sim.info.tobytes()
    node.close()


def test_cuda_graph_replay_equals_single_launches(orc):
    nx, ny = 640, 200
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    a = sb.D2Q9Node((nx * 2, ny * 2), setting(W.POISEUILLE), lattice=(nx, ny), lattice_info=info)
    b = sb.D2Q9Node((nx * 2, ny * 2), setting(W.POISEUILLE), lattice=(nx, ny), lattice_info=info, flags=sb.FLAG_NO_GRAPH)
    for n in (1, 33, 70, 16, 5):  # >= 32 updates go through 16-update graphs, the rest one by one
        a.step_n(n)
        b.step_n(n)
        assert a.swap_index == b.swap_index
    assert a.launch_count == b.launch_count
    for which in (0, 1):
        assert_bits_equal(a.read_distributions(which), b.read_distributions(which), f"graph vs launches buf{which}")
    # a mask change rebuilds the graphs
    patch = info[100 * nx:101 * nx].copy()
    patch["material"][300:320] = W.OBSTACLE
    for n in (a, b):
        n.write_lattice_info(100 * nx * 16, patch)
        n.step_n(40)
    sim = oracle_for(orc, nx, ny, info)
    sim.step(125)
    sim.write_lattice_info(100 * nx * 16, patch)
    sim.step(40)
    live = np.ones((ny, nx), bool)
    live[100, 300:320] = False
    for n in (a, b):
        got = n.read_distributions(n.swap_index)
        assert_bits_equal(got[:, live], sim.distributions(sim.swap)[:, live], "after mask change")
    a.close()
    b.close()


def test_frames_graph_equals_python_frame_loop(orc):
    canvas = (400, 260)
    fa = sb.FluidSimulator(canvas, setting(W.POISEUILLE), particles=True)
    fb = sb.FluidSimulator(canvas, setting(W.POISEUILLE), particles=True, flags=sb.FLAG_NO_GRAPH)
    fa.compute(25)  # one call, graph replay
    for _ in range(25):  # the reference's frame loop spelled out (fluid_simulator.rs:223-231)
        nb = fb.fluid_compute_node
        nb.compute_by_pass(0)
        nb.particles_update()
        nb.compute_by_pass(1)
        nb.particles_update()
    na, nb = fa.fluid_compute_node, fb.fluid_compute_node
    n = fa.particles_num[0] * fa.particles_num[1]
    assert na.read_particles(n).tobytes() == nb.read_particles(n).tobytes()
    assert_bits_equal(na.read_distributions(0), nb.read_distributions(0), "frames")
    assert na.swap_index == nb.swap_index == 0
    na.close()
    nb.close()


def test_checkpoint_restore_roundtrip(orc):
    """lbm_read_* / lbm_write_* restore a run exactly (the reference cannot read its buffers back)."""
    nx, ny = 300, 140
    info = orc.init_lattice_material(nx, ny, W.LID_DRIVEN_CAVITY)
    a = sb.D2Q9Node((nx * 2, ny * 2), setting(W.LID_DRIVEN_CAVITY), lattice=(nx, ny), lattice_info=info)
    a.step_n(51)
    snap = [a.read_distributions(0), a.read_distributions(1)]
    snap_info, swap = a.read_lattice_info(), a.swap_index
    a.step_n(40)
    want = a.read_distributions(a.swap_index)
    b = sb.D2Q9Node((nx * 2, ny * 2), setting(W.LID_DRIVEN_CAVITY), lattice=(nx, ny), lattice_info=snap_info)
    b.write_distributions(0, snap[0])
    b.write_distributions(1, snap[1])
    for k in range(40):  # resume with the explicit swap index, like compute_by_pass
        b.compute_by_pass((swap + k) % 2)
    assert_bits_equal(b.read_distributions(b.swap_index), want, "resumed run")
    assert_bits_equal(b.read_distributions(0), b.read_distributions(0), "self")
    a.close()
    b.close()


def test_partial_and_out_of_range_lattice_info_writes(orc):
    nx, ny = 64, 40
    info = orc.init_lattice_material(nx, ny, W.CUSTOM)
    node = sb.D2Q9Node((nx * 2, ny * 2), setting(W.CUSTOM), lattice=(nx, ny), lattice_info=info)
    h = node._h
    total = nx * ny * 16
    one = np.zeros(1, W.LATTICE_INFO_DTYPE)
    one[0] = (W.OBSTACLE, -1, 0.0, 0.0)
    assert lib.lbm_write_lattice_info(h, total, ptr(one), 16) == _capi.ERR_INVALID_ARG      # past the end
    assert lib.lbm_write_lattice_info(h, total - 8, ptr(one), 16) == _capi.ERR_INVALID_ARG  # straddles the end
    assert lib.lbm_write_lattice_info(h, 0, None, 16) == _capi.ERR_INVALID_ARG
    assert lib.lbm_write_lattice_info(h, 160, ptr(one), 0) == _capi.OK                     # empty write
    # a 4-byte write of just the material field of one cell (any byte range is accepted)
    mat = np.array([W.OBSTACLE], np.int32)
    off = (20 * nx + 30) * 16
    assert lib.lbm_write_lattice_info(h, off, ptr(mat), 4) == _capi.OK
    got = node.read_lattice_info().reshape(ny, nx)
    assert got["material"][20, 30] == W.OBSTACLE and got["block_iter"][20, 30] == -1
    want = info.copy()
    want.reshape(ny, nx)["material"][20, 30] = W.OBSTACLE
    assert got.tobytes() == want.tobytes()
    sim = oracle_for(orc, nx, ny, info)
    sim.write_lattice_info(off, want[20 * nx + 30:20 * nx + 31])
    node.step_n(30)
    sim.step(30)
    live = np.ones((ny, nx), bool)
    live[20, 30] = False
    assert_bits_equal(node.read_distributions(node.swap_index)[:, live], sim.distributions(sim.swap)[:, live], "live")
    node.close()


def test_call_order_errors_are_reported_not_fatal():
    d = _capi.LbmDesc()
    d.struct_size = C.sizeof(_capi.LbmDesc)
    d.nx, d.ny, d.lattice_pixel_size, d.device, d.world = 64, 48, 2, -1, 1
    h = C.c_void_p()
    assert lib.lbm_create(C.byref(d), C.byref(h)) == _capi.OK
    assert lib.lbm_step(h, 0) == _capi.ERR_STATE and b"lbm_write_uniform" in lib.lbm_last_error(h)
    u = sb.lbm_uniform_new(0.56, 0, 64 * 48)
    assert lib.lbm_write_uniform(h, C.byref(u)) == _capi.OK
    assert lib.lbm_reset(h) == _capi.ERR_STATE and b"lattice info" in lib.lbm_last_error(h)
    assert lib.lbm_generate_lattice_info(h, 99, 0, 0.0) == _capi.ERR_INVALID_ARG
    assert lib.lbm_generate_lattice_info(h, W.POISEUILLE, 0, 0.0) == _capi.OK
    assert lib.lbm_reset(h) == _capi.OK and lib.lbm_step_n(h, 4) == _capi.OK
    assert lib.lbm_particles_update(h) == _capi.ERR_STATE      # created without particles
    assert lib.lbm_canvas_clear(h) == _capi.ERR_STATE
    blob = _capi.LbmIpcBlob()
    assert lib.lbm_ipc_export(h, C.byref(blob)) == _capi.OK
    assert lib.lbm_ipc_attach(h, C.byref(blob), C.byref(blob)) == _capi.ERR_STATE  # single slab has no neighbours
    f = W.FieldUniform()
    lib.lbm_field_uniform_new(32, 32, 2, 64, 64, C.byref(f))
    assert lib.lbm_write_field_uniform(h, C.byref(f)) == _capi.ERR_INVALID_ARG     # wrong lattice size
    ms = C.c_float()
    assert lib.lbm_last_step_n_ms(h, C.byref(ms)) == _capi.OK and ms.value > 0
    assert lib.lbm_sync(h) == _capi.OK
    lib.lbm_destroy(h)
    # a slab that is not attached refuses to step
    d.rank, d.world = 0, 2
    assert lib.lbm_create(C.byref(d), C.byref(h)) == _capi.OK
    lib.lbm_write_uniform(h, C.byref(u))
    lib.lbm_generate_lattice_info(h, W.POISEUILLE, 0, 0.0)
    assert lib.lbm_step_n(h, 1) == _capi.ERR_STATE and b"lbm_ipc_attach" in lib.lbm_last_error(h)
    lib.lbm_destroy(h)


def test_lost_neighbour_times_out_instead_of_hanging(orc):
    """A slab whose neighbour is never stepped must not hang the GPU: the edge CTAs give up after a
    bounded wait and lbm_sync reports it."""
    import time

    from simuverse_b200.slabs import SlabGroup

    nx, ny = 128, 64
    info = orc.init_lattice_material(nx, ny, W.CUSTOM)
    grp = SlabGroup((nx * 2, ny * 2), setting(W.CUSTOM), lattice=(nx, ny), n_slabs=2, lattice_info=info)
    a, b = grp.nodes
    a.step_n(1)           # fine: neighbours' progress (0) is enough for the first update
    a.sync()
    t0 = time.time()
    a.step_n(1)           # needs slab b's first update, which never comes
    with pytest.raises(sb.LbmError) as e:
        a.sync()
    assert e.value.status == _capi.ERR_STATE and "neighbour" in str(e.value)
    assert 2.0 < time.time() - t0 < 30.0
    b.sync()
    a.close()
    b.close()


def test_pipelined_macro_readback(orc):
    """lbm_read_macro_async: the field of frame k is copied out while frame k+1 is computed into a second
    texture; every delivered field equals the oracle's texture of that frame."""
    nx, ny = 256, 160
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    node = sb.D2Q9Node((nx * 2, ny * 2), setting(W.POISEUILLE), lattice=(nx, ny), lattice_info=info,
                       flags=sb.FLAG_MACRO_EVERY_STEP)
    sim = oracle_for(orc, nx, ny, info)
    outs = [np.zeros((ny, nx, 4), np.float16) for _ in range(5)]
    want = []
    for k in range(5):
        node.compute_frames(1)
        node.read_macro_tex_async(outs[k])
        sim.step(2)
        want.append(sim.macro_f16.copy())
    # right after the texture switch the synchronous read must still return the newest field
    np.testing.assert_array_equal(node.read_macro_tex().view(np.uint16).reshape(-1), want[-1])
    node.sync()
    for k in range(5):
        np.testing.assert_array_equal(outs[k].view(np.uint16).reshape(-1), want[k])
    node.step_n(40)  # graphs are rebuilt on the new texture
    sim.step(40)
    np.testing.assert_array_equal(node.read_macro_tex().view(np.uint16).reshape(-1), sim.macro_f16)
    for which in (0, 1):
        assert_bits_equal(node.read_distributions(which), sim.distributions(which), "state")
    node.close()


def test_cpp_host_mirror_on_the_gpu(tmp_path):
    """include/d2q9_node.hpp (the C++ mirror of D2Q9Node / FluidSimulator, the language a compiled host would use in
    place of the unbuildable Rust shim) through a session of frames, a click and a drag: frames issued through
    lbm_compute_frames (sweeps) and call by call (single updates) leave identical distributions."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host_mirror_check"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "host_mirror_check.cpp"), sb.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(sb.LIB_PATH), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOST_MIRROR_OK gpu" in r.stdout, r.stdout + r.stderr
