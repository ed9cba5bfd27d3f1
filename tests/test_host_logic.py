"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol
include/lbm_b200.h declares, the host mirrors of the reference's Rust helpers agree with the
oracle bit for bit, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import simuverse_b200 as sb
from simuverse_b200 import _capi
from simuverse_b200 import wire as W
from simuverse_b200.wire import ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 40
    raw = C.CDLL(sb.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/lbm_b200.h but not exported"
        assert n in _capi.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_capi.PROTOTYPES) == names
    assert sb.lib.lbm_abi_version() == 1


def test_wire_sizes_match_header():
    text = open(os.path.join(ROOT, "include", "lbm_wire.h")).read()
    for name, size in [("LbmUniform", 304), ("FieldUniform", 48), ("LatticeInfo", 16), ("ParticleUniform", 48),
                       ("TrajectoryParticle", 24), ("Pixel", 12)]:
        assert f"sizeof({name}) == {size}" in text
        assert C.sizeof(getattr(W, name)) == size
    assert C.sizeof(_capi.LbmDesc) == 44 and C.sizeof(_capi.LbmIpcBlob) == 256


def test_uniform_new_bytes_match_oracle(orc):
    for tau, ty in [(0.56, 0), (0.8, 1), (1.7, 0)]:
        a = sb.lbm_uniform_new(tau, ty, 225000)
        b = orc.uniform_new(tau, ty, 225000)
        assert bytes(a) == bytes(b)
    assert sb.lib.lbm_tau_from_viscosity(float(np.float32(0.02))) == orc.tau_from_viscosity(0.02)


@pytest.mark.parametrize("nx,ny", [(600, 375), (128, 128), (131, 77), (37, 19), (300, 1000)])
@pytest.mark.parametrize("ty", [W.POISEUILLE, W.LID_DRIVEN_CAVITY, W.CUSTOM, 0])
def test_init_lattice_material_matches_oracle(orc, nx, ny, ty):
    a = sb.init_lattice_material(nx, ny, ty)
    b = orc.init_lattice_material(nx, ny, ty)
    assert a.tobytes() == b.tobytes()


def test_porous_material_matches_oracle(orc):
    a = sb.init_porous_material(257, 130, seed=0x5EED, solid_fraction=0.30)
    b = orc.init_porous_material(257, 130, seed=0x5EED, solid_fraction=0.30)
    assert a.tobytes() == b.tobytes()
    inner = a.reshape(130, 257)["material"][1:-1, 2:-2]
    frac = (inner == W.OBSTACLE).mean()
    assert 0.27 < frac < 0.33
    c = sb.init_porous_material(257, 130, seed=7, solid_fraction=0.30)
    assert c.tobytes() != a.tobytes()


def test_on_click_guard_and_obstacle_patch_match_oracle(orc):
    nx, ny, lps = 600, 375, 2
    for pos in [(0.0, 10.0), (-1.0, 5.0), (55.9, 300.0), (56.0, 56.0), (600.0, 400.0), (1139.0, 689.0), (1140.0, 400.0),
                (700.5, 690.0)]:
        x, y = C.c_uint32(), C.c_uint32()
        ok = sb.lib.lbm_on_click_guard(nx, ny, lps, pos[0], pos[1], C.byref(x), C.byref(y))
        want = orc.on_click_guard(nx, ny, lps, *pos)
        assert (ok == 1) == (want is not None)
        if want:
            assert (x.value, y.value) == want
    mirror_a = sb.init_lattice_material(nx, ny, W.POISEUILLE)
    mirror_b = mirror_a.copy()
    for (x, y) in [(300, 200), (305, 190), (40, 28), (569, 344)]:
        patch = np.zeros(56 * nx, W.LATTICE_INFO_DTYPE)
        off = C.c_uint64()
        n = sb.lib.lbm_obstacle_patch(nx, ny, ptr(mirror_a), x, y, ptr(patch), C.byref(off))
        woff, wpatch = orc.add_obstacle(nx, ny, mirror_b, x, y)
        assert n == wpatch.size == 56 * nx and off.value == woff == nx * (y - 28) * 16
        assert patch[:n].tobytes() == wpatch.tobytes()
        assert mirror_a.tobytes() == mirror_b.tobytes()
    assert (mirror_a["material"] == W.OBSTACLE).sum() > 7374 + 3 * 2400


def test_external_force_cells_match_oracle(orc):
    nx, ny, lps = 600, 375, 2
    cases = [((400.0, 300.0), (380.0, 310.0)), ((10.0, 10.0), (300.0, 300.0)), ((500.5, 200.25), (500.5, 200.25)),
             ((1198.0, 700.0), (1100.0, 745.0)), ((3.0, 3.0), (1.0, 2.0))]
    for pos, pre in cases:
        offs = np.zeros(4096, np.uint64)
        cells = np.zeros(4096, W.LATTICE_INFO_DTYPE)
        n = sb.lib.lbm_external_force_cells(nx, ny, lps, pos[0], pos[1], pre[0], pre[1], ptr(offs), ptr(cells), 4096)
        woffs, wcells = orc.add_external_force(nx, ny, lps, pos, pre)
        assert n == woffs.size
        np.testing.assert_array_equal(offs[:n], woffs)
        assert cells[:n].tobytes() == wcells.tobytes()
        if n:
            assert (cells[:n]["material"] == 6).all() and (cells[:n]["block_iter"] == 90).all()
            assert np.hypot(cells[0]["vx"], cells[0]["vy"]) <= 0.12 + 1e-6


def test_particle_seeding_matches_oracle(orc):
    assert sb.particle_grid((1200, 750), 10000) == orc.particle_grid(1200, 750, 10000) == (127, 80)
    assert sb.particle_grid((16384, 16384), 1000000) == orc.particle_grid(16384, 16384, 1000000) == (1000, 1000)
    for life in (90.0, 1.0, 0.0):
        a = sb.init_trajectory_particles((1200, 750), (127, 80), life, 0x5EED)
        b = orc.init_trajectory_particles(1200, 750, 127, 80, life, 0x5EED)
        assert a.tobytes() == b.tobytes()
    a = sb.init_trajectory_particles((1200, 750), (127, 80), 90.0, 0x5EED)
    assert (a["life_time"] >= 0).all() and (a["life_time"] <= 90).all() and (a["fade"] == 0).all()
    np.testing.assert_array_equal(a["pos"], a["pos_initial"])


def test_field_uniform_matches_oracle(orc):
    f = W.FieldUniform()
    sb.lib.lbm_field_uniform_new(600, 375, 2, 1200, 750, C.byref(f))
    assert bytes(f) == bytes(orc.field_uniform_new(600, 375, 2, 1200, 750))
    assert f.speed_ty == 1 and tuple(f.lattice_pixel_size) == (2.0, 2.0)


def test_no_gpu_means_loud_failure_not_fallback():
    """On a box without a CUDA device the compute path must refuse to exist."""
    if sb.lib.lbm_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sb.LbmError) as e:
        sb.D2Q9Node((1200, 750), sb.SettingObj())
    assert e.value.status == _capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_create_rejects_bad_descriptors():
    d = _capi.LbmDesc()
    h = C.c_void_p()
    d.struct_size = 4
    assert sb.lib.lbm_create(C.byref(d), C.byref(h)) == _capi.ERR_INVALID_ARG
    d.struct_size = C.sizeof(_capi.LbmDesc)
    d.nx, d.ny, d.lattice_pixel_size, d.world = 2, 2, 2, 1
    assert sb.lib.lbm_create(C.byref(d), C.byref(h)) == _capi.ERR_INVALID_ARG
    assert b"too small" in sb.lib.lbm_last_error(None)
    assert sb.lib.lbm_create(None, C.byref(h)) == _capi.ERR_INVALID_ARG
    assert sb.lib.lbm_step(None, 0) == _capi.ERR_INVALID_ARG  # null handle never crashes


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "simuverse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_reference_arm_never_maps_the_product_library():
    """`bench.py --impl reference` (and anything else that only needs the oracle) must not load liblbm_b200.so:
    tests/oracle.py loads wire.py on its own when the product package has not been imported."""
    import json
    import subprocess
    import sys

    code = ("import sys, os; sys.path.insert(0, %r); import oracle; oracle.lib(); "
            "assert not any(m.startswith('simuverse_b200') for m in sys.modules), 'package imported'; "
            "maps = open('/proc/self/maps').read(); assert 'liblbm_b200' not in maps, 'product library mapped'; "
            "assert 'liblbm_oracle' in maps; print('CLEAN')" % os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "CLEAN" in r.stdout, r.stdout + r.stderr
    # the arm itself, on a small lattice: one JSON line whose `config` has the keys of the b200 arm's
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                        "--lattice", "256", "128"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port"
    assert set(line["config"]) == {"workload", "baseline_config", "lattice", "tau", "l2"}
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_cpp_host_mirror_compiles_and_runs(tmp_path):
    """include/d2q9_node.hpp (C++ mirror of D2Q9Node / FluidSimulator) builds against the library;
    on a CPU box it must report the missing GPU instead of computing anything."""
    import subprocess

    exe = tmp_path / "host_mirror_check"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_check.cpp"), sb.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(sb.LIB_PATH), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "HOST_MIRROR_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("h,rows", [(4096, 16), (2048, 31), (45, 1), (45, 2), (45, 64), (33, 8), (4, 8), (17, 16), (5, 2)])
def test_sweep_row_blocks_cover_every_row_once(h, rows):
    """Work items of the two-update sweep (lbm_sweep_blocks): a partition of [0, h) in dispatch order with the
    blocks holding the first and the last rows up front, and a last block of at least two rows."""
    import ctypes as C

    from simuverse_b200._capi import lib

    cap = h // rows + 2
    out = (C.c_int32 * (2 * cap))()
    n_edge = C.c_int32(0)
    n = lib.lbm_sweep_blocks(h, rows, out, cap, C.byref(n_edge))
    assert n >= 1
    blocks = [(out[2 * k], out[2 * k + 1]) for k in range(n)]
    covered = sorted(blocks)
    assert covered[0][0] == 0 and covered[-1][1] == h
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:])), "blocks must tile the rows without gaps"
    assert all(y1 > y0 for y0, y1 in blocks)
    assert blocks[0][0] == 0
    assert n_edge.value == (2 if n > 1 else 1)
    if n > 1:
        assert blocks[1][1] == h and blocks[1][1] - blocks[1][0] >= 2 or h < 2
        assert [b for b in blocks[2:]] == covered[1:-1], "interior blocks top to bottom"
    assert lib.lbm_sweep_blocks(h, rows, out, 0, C.byref(n_edge)) == 0  # cap too small: refused, nothing written


@pytest.mark.parametrize("h,rows,tail,tail_rows", [(4096, 16, 352, 8), (2048, 32, 192, 16), (100, 8, 24, 4), (64, 16, 0, 8),
                                                   (64, 16, 64, 8), (33, 8, 9, 2)])
def test_sweep_row_blocks_with_a_short_tail(h, rows, tail, tail_rows):
    """lbm_sweep_blocks_tail: same partition rules; the last `tail` rows come in shorter blocks, dispatched last."""
    import ctypes as C

    from simuverse_b200._capi import lib

    cap = h // min(rows, tail_rows) + 4
    out = (C.c_int32 * (2 * cap))()
    n_edge = C.c_int32(0)
    n = lib.lbm_sweep_blocks_tail(h, rows, tail, tail_rows, out, cap, C.byref(n_edge))
    assert n >= 1
    blocks = [(out[2 * k], out[2 * k + 1]) for k in range(n)]
    covered = sorted(blocks)
    assert covered[0][0] == 0 and covered[-1][1] == h
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    assert blocks[0][0] == 0 and (n == 1 or blocks[1][1] == h)
    assert [b for b in blocks[2:]] == covered[1:-1]
    eff_tail = tail if 0 <= tail < h else 0
    for y0, y1 in covered:
        limit = tail_rows if y0 >= h - eff_tail else rows
        assert y1 - y0 <= limit + 1  # (+1: a 1-row remainder is merged into its predecessor)
    if eff_tail:
        assert any(y0 >= h - eff_tail and y1 - y0 <= tail_rows for y0, y1 in covered)


def test_schedule_verdict_of_a_mask_write_is_a_function_of_its_bytes():
    """lbm_scan_lattice_info_write: what every rank of a multi-slab lattice derives from a lbm_write_lattice_info call."""
    import ctypes as C

    from simuverse_b200 import wire as W
    from simuverse_b200._capi import lib

    nx, ny = 64, 40

    def scan(off, cells):
        cells = np.ascontiguousarray(cells, W.LATTICE_INFO_DTYPE)
        a, b = C.c_int32(-7), C.c_int32(-7)
        ok = lib.lbm_scan_lattice_info_write(nx, ny, off, W.ptr(cells), cells.nbytes, C.byref(a), C.byref(b))
        return ok, a.value, b.value

    bulk = np.zeros(3 * nx, W.LATTICE_INFO_DTYPE)
    bulk["material"], bulk["block_iter"] = W.BULK, -1
    assert scan(10 * nx * 16, bulk) == (1, 0, 0)                       # three interior rows of bulk
    obst = bulk.copy()
    obst["material"][nx + 20:nx + 30] = W.OBSTACLE                      # an obstacle in the interior
    assert scan(10 * nx * 16, obst) == (1, 0, 0)
    assert scan(0, obst) == (1, 0, 1)                                   # the same bytes at rows 0..2: row 1 is next to the ring
    edge = bulk.copy()
    edge["material"][nx + 1] = W.OBSTACLE                               # x = 1 of the second written row
    assert scan(10 * nx * 16, edge) == (1, 0, 1)
    edge = bulk.copy()
    edge["material"][2 * nx - 2] = W.BOUNDARY                           # x = nx - 2
    assert scan(10 * nx * 16, edge) == (1, 0, 1)
    one = np.zeros(1, W.LATTICE_INFO_DTYPE)
    one[0] = (W.EXTERNAL_FORCE, 90, 0.01, 0.0)                          # add_external_force: one armed cell
    assert scan((20 * nx + 33) * 16, one) == (1, 90, 0)
    one[0] = (W.EXTERNAL_FORCE, -1, 0.01, 0.0)                          # permanent force cell: nothing counts down
    assert scan((20 * nx + 33) * 16, one) == (1, 0, 0)
    one[0] = (W.OBSTACLE, -1, 0.0, 0.0)
    assert scan(((ny - 2) * nx + 33) * 16, one) == (1, 0, 1)            # y = ny - 2
    # a run that wraps from the end of one row to the start of the next: x restarts at 0
    two = np.zeros(2, W.LATTICE_INFO_DTYPE)
    two[:] = (W.BULK, -1, 0.0, 0.0)
    two[1] = (W.OBSTACLE, -1, 0.0, 0.0)
    assert scan((20 * nx + nx - 1) * 16, two) == (1, 0, 1)              # second cell is (x = 0, y = 21)
    assert scan((20 * nx + 30) * 16 + 4, one)[0] == 0                   # misaligned: refused ...
    assert scan((20 * nx + 30) * 16 + 4, one)[2] == 1                   # ... with the conservative verdict
