"""Pins the CPU oracle: derived known answers, the committed golden vectors,
and bit-agreement with the independent numpy restatement.  CPU only."""
import ctypes as C
import zlib

import numpy as np
import pytest

import np_restatement as R
from helpers import WGSL_DEFAULT, assert_bits_equal, golden_cases, sha, tau_default, wgsl_golden_cases
from simuverse_b200 import wire as W


def test_uniform_constants(orc):
    u = orc.uniform_new(tau_default(), 0, 600 * 375)
    assert C.sizeof(u) == 304
    assert abs(u.tau - 0.56) < 1e-7 and u.omega == np.float32(1.0) / np.float32(u.tau)
    w = [u.e_w_max[i][2] for i in range(9)]
    assert w == [np.float32(0.444444)] + [np.float32(0.111111)] * 4 + [np.float32(0.0277777)] * 4
    assert [u.e_w_max[i][3] for i in range(9)] == [np.float32(0.6)] + [np.float32(0.2222)] * 4 + [np.float32(0.1111)] * 4
    assert [u.inversed_direction[i][0] for i in range(9)] == [0, 3, 4, 1, 2, 7, 8, 5, 6]
    e = [(u.e_w_max[i][0], u.e_w_max[i][1]) for i in range(9)]
    assert e == [(0, 0), (1, 0), (0, -1), (-1, 0), (0, 1), (1, -1), (-1, -1), (-1, 1), (1, 1)]
    for i in range(9):  # inverse really is the opposite vector
        j = u.inversed_direction[i][0]
        assert (e[j][0], e[j][1]) == (-e[i][0], -e[i][1])


@pytest.mark.parametrize("nx,ny,hist,crc", [
    (600, 375, {1: 214934, 2: 1200, 3: 373, 4: 7374, 5: 373, 7: 746}, 0xF8EFC262),
    (4096, 4096, {1: 16745276, 2: 8192, 3: 4094, 4: 7372, 5: 4094, 7: 8188}, 0x63A17811),
])
def test_mask_known_answers(orc, nx, ny, hist, crc):
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    m = info["material"]
    got = {int(k): int(v) for k, v in zip(*np.unique(m, return_counts=True))}
    assert got == hist
    assert zlib.crc32(m.astype("<i4").tobytes()) == crc
    assert (info["block_iter"] == -1).all() and (info["vy"] == 0).all()
    assert (info["vx"][m == 3] == np.float32(0.12)).all() and (info["vx"][m != 3] == 0).all()


def test_init_known_answer(orc):
    nx, ny = 64, 40
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    s = orc.OracleSim(nx, ny, info, orc.uniform_new(tau_default(), 0, nx * ny))
    b0, b1 = s.distributions(0), s.distributions(1)
    fluid = [0.444444, 0.1666665, 0.111111, 0.0555555, 0.111111, 0.0277777, 0.0277777, 0.0277777, 0.0277777]
    # (50, 20) is bulk fluid: the R=28 preset discs sit at x <= 41 on a 64-wide lattice
    np.testing.assert_allclose(b0[:, 20, 50], np.array(fluid, np.float32), rtol=0, atol=1e-7)
    assert b0[1, 20, 50] == np.float32(0.111111) + np.float32(0.111111) * np.float32(0.5)
    np.testing.assert_array_equal(b1[:, 20, 50], [0, b0[1, 20, 50], 0, b0[3, 20, 50], 0, 0, 0, 0, 0])
    assert (b0[:, 0, :] == 0).all() and (b1[:, 0, :] == 0).all()  # wall row
    # lid-driven cavity: no +x bias
    info = orc.init_lattice_material(nx, ny, W.LID_DRIVEN_CAVITY)
    s = orc.OracleSim(nx, ny, info, orc.uniform_new(tau_default(), 1, nx * ny))
    assert s.distributions(0)[1, 5, 5] == np.float32(0.111111) and (s.distributions(1)[:, 5, 5] == 0).all()


def test_total_mass_known_answers(orc):
    """M0/M1/M100 of the default 600x375 channel, tau=0.56."""
    nx, ny = 600, 375
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    s = orc.OracleSim(nx, ny, info, orc.uniform_new(tau_default(), 0, nx * ny), threads=orc.lib().orc_get_max_threads())
    assert round(s.total_mass(), 3) == 216425.742
    s.step(1)
    assert round(s.total_mass(), 3) == 216148.638
    s.step(99)
    assert round(s.total_mass(), 3) == 216071.546


def test_uniform_periodic_cell_hand_computed(orc):
    """One step of an all-fluid periodic lattice from f=w: rho = sum w (clamped no-op), u = 0,
    feq_i = rho*w_i*1.0, t = f - omega*(f - feq)."""
    nx, ny = 8, 6
    info = np.zeros(nx * ny, W.LATTICE_INFO_DTYPE)
    info["material"] = W.BULK
    info["block_iter"] = -1
    u = orc.uniform_new(tau_default(), 1, nx * ny)
    s = orc.OracleSim(nx, ny, info, u)
    s.step(1)
    w = np.array([u.e_w_max[i][2] for i in range(9)], np.float32)
    rho = np.float32(0)
    for i in range(9):
        rho = np.float32(rho + w[i])
    om = np.float32(u.omega)
    for i in range(9):
        feq = np.float32(np.float32(rho * w[i]) * np.float32(1.0))
        t = np.float32(w[i] - np.float32(om * np.float32(w[i] - feq)))
        assert (s.distributions(1)[i] == t).all()
    m = s.macro()
    assert (m[0] == 0).all() and (m[1] == 0).all() and (m[2] == rho).all()


def test_bounce_back_known_answer(orc):
    """A fluid cell next to a wall gets its own post-collision f*_i back in inv(i) one step later;
    the value sits in the solid neighbour's slot and the fluid slot is zeroed (boundary.wgsl:28-31).
    Ring cells (ghost column) receive nothing from a solid neighbour."""
    nx, ny = 12, 9
    info = orc.init_lattice_material(nx, ny, W.CUSTOM)
    u = orc.uniform_new(tau_default(), 1, nx * ny)
    s = orc.OracleSim(nx, ny, info, u)
    before = s.distributions(0).copy()
    s.step(1)
    after = s.distributions(1)
    # cell (x=1,y=4) is strictly interior with the wall x=0 to its left: direction 3 (e=(-1,0))
    x, y = 1, 4
    assert after[3, y, x] == 0.0                     # own slot zeroed
    assert after[1, y, x - 1] != 0.0                 # f*_3 parked in wall slot inv(3)=1
    s.step(1)
    # the second step pulled direction 1 at (1,4) from the wall slot = its own f*_3
    prev = s.distributions(1)
    assert prev[1, y, x - 1] != 0 and before[3, y, x] != 0
    # first step after init pulled zeros from solids: rho at wall-adjacent cells hit the 0.8 clamp
    info2 = orc.init_lattice_material(64, 40, W.POISEUILLE)
    s2 = orc.OracleSim(64, 40, info2, orc.uniform_new(tau_default(), 0, 64 * 40))
    s2.step(1)
    assert (s2.macro()[2][1, 2:-2] == np.float32(0.8)).any()
    # ghost cell (x=0, y=1) neighbours the wall row y=0 but is on the ring: its slots facing the wall keep values
    d = s2.distributions(1)
    assert d[2, 1, 0] != 0.0 and d[4, 0, 0] == 0.0          # the ring cell (0,1) parks nothing in the wall above it
    assert d[4, 0, 50] != 0.0 and d[2, 1, 50] == 0.0        # interior cell (50,1): f*_2 parked in wall slot 4


def test_transient_force_cell_flips_after_block_iter(orc):
    nx, ny = 20, 16
    info = orc.init_lattice_material(nx, ny, W.CUSTOM)
    u = orc.uniform_new(tau_default(), 1, nx * ny)
    s = orc.OracleSim(nx, ny, info, u)  # init first (it would disarm block_iter>0 cells)
    cell = np.zeros(1, W.LATTICE_INFO_DTYPE)
    cell[0] = (W.EXTERNAL_FORCE, 5, 0.05, -0.02)
    s.write_lattice_info((nx * 7 + 9) * 16, cell)
    for k in range(1, 8):
        s.step(1)
        c = s.info[nx * 7 + 9]
        if k < 5:
            assert (c["material"], c["block_iter"]) == (6, 5 - k)
        else:
            assert (c["material"], c["block_iter"]) == (1, 0)
        assert c["vx"] == np.float32(0.05)  # vx/vy are retained after the flip
    # init.wgsl disarms armed cells
    s.write_lattice_info((nx * 7 + 9) * 16, cell)
    s.reset()
    c = s.info[nx * 7 + 9]
    assert (c["material"], c["block_iter"], c["vx"], c["vy"]) == (1, 0, 0.0, 0.0)


def test_direction_clamp_trips(orc):
    """Per-direction clamp [0, max_i] (collide_stream.wgsl:79-83) trips next to a strong force patch."""
    nx, ny = 24, 20
    info = orc.init_lattice_material(nx, ny, W.CUSTOM)
    g = info.reshape(ny, nx)
    g[8:12, 8:12] = (W.EXTERNAL_FORCE, -1, 0.9, 0.9)
    u = orc.uniform_new(tau_default(), 1, nx * ny)
    s = orc.OracleSim(nx, ny, info, u)
    s.step(3)
    cur = s.distributions(s.swap)
    mx = np.array([u.e_w_max[i][3] for i in range(9)], np.float32)
    assert (cur >= 0).all() and all((cur[i] <= mx[i]).all() for i in range(9))
    assert any((cur[i] == mx[i]).any() for i in range(9)) and (cur[:, 8:12, 8:12] == 0).any()


@pytest.mark.parametrize("path", golden_cases())
def test_oracle_matches_golden(orc, path):
    g = np.load(path)
    nx, ny, steps = int(g["nx"]), int(g["ny"]), int(g["steps"])
    fluid_ty = 1 if int(g["preset"]) == W.LID_DRIVEN_CAVITY else 0
    s = orc.OracleSim(nx, ny, g["info"], orc.uniform_new(tau_default(), fluid_ty, nx * ny))
    for off, cell in zip(g["post_offsets"], g["post_cells"]):  # force cells armed after init
        s.write_lattice_info(int(off), np.array([cell], W.LATTICE_INFO_DTYPE))
    s.step(steps)
    assert s.swap == int(g["swap"])
    assert_bits_equal(s.distributions(s.swap), g["buf_cur"], "current buffer")
    assert_bits_equal(s.distributions(1 - s.swap), g["buf_prev"], "previous buffer")
    m = s.macro()
    assert_bits_equal(m[0], g["ux"], "ux")
    assert_bits_equal(m[1], g["uy"], "uy")
    assert_bits_equal(m[2], g["rho"], "rho")
    np.testing.assert_array_equal(s.info["material"].reshape(ny, nx), g["material"])
    np.testing.assert_array_equal(s.info["block_iter"].reshape(ny, nx), g["block_iter"])


@pytest.mark.parametrize("nx,ny,preset,steps", [(150, 94, W.POISEUILLE, 120), (33, 27, W.LID_DRIVEN_CAVITY, 60),
                                                (17, 9, W.CUSTOM, 40)])
def test_oracle_matches_numpy_restatement(orc, nx, ny, preset, steps):
    info = orc.init_lattice_material(nx, ny, preset)
    fluid_ty = 1 if preset == W.LID_DRIVEN_CAVITY else 0
    u = orc.uniform_new(tau_default(), fluid_ty, nx * ny)
    a = orc.OracleSim(nx, ny, info, u, threads=3)  # OpenMP on: collide pass is order independent
    b = R.NpSim(nx, ny, info, u)
    for _ in range(2):
        a.step(steps // 2)
        b.step(steps // 2)
        assert_bits_equal(a.distributions(a.swap), b.buf[b.swap], "current")
        assert_bits_equal(a.distributions(1 - a.swap), b.buf[1 - b.swap], "previous")
    assert_bits_equal(a.macro(), np.stack(b.macro), "macro")


def test_f16_conversion_matches_numpy(orc):
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.standard_normal(4000).astype(np.float32) * np.float32(0.2),
        np.array([0.0, -0.0, 1.0, 0.8, 1.2, 65504.0, 65519.9, 65520.0, 1e-8, 6e-8, 2.98e-8, 2.9802322e-8, 6.1e-5, 6.0975e-5,
                  -0.12, 1e10, np.inf, -np.inf], np.float32),
        np.float32(2.0) ** rng.integers(-26, 16, 500).astype(np.float32) * (1 + rng.random(500).astype(np.float32)),
    ])
    got = orc.f32_to_f16_bits(vals)
    want = vals.astype(np.float16).view(np.uint16)
    np.testing.assert_array_equal(got, want)
    back = np.array([orc.lib().orc_f16_to_f32(int(h)) for h in got[:200]], np.float32)
    np.testing.assert_array_equal(back.view(np.uint32), want[:200].view(np.float16).astype(np.float32).view(np.uint32))


def test_mass_drifts_like_the_reference(orc):
    """The reference does not conserve mass (truncated weights + clamps): the oracle must
    show the drift rather than hide it."""
    nx, ny = 128, 128
    info = orc.init_lattice_material(nx, ny, W.CUSTOM)
    s = orc.OracleSim(nx, ny, info, orc.uniform_new(tau_default(), 1, nx * ny), threads=4)
    m0 = s.total_mass()
    s.step(400)
    drift = (s.total_mass() - m0) / m0
    assert -5e-2 < drift < -1e-5


# ------------------------------------------------------------------ pinned to the reference's WGSL source

@pytest.mark.parametrize("path", wgsl_golden_cases())
def test_oracle_matches_executed_reference_wgsl(orc, path):
    """The golden vectors come from the reference's own shader text, transpiled and executed
    (tests/golden/make_wgsl_golden.py): init, collide_stream, boundary and particle_update."""
    g = np.load(path)
    nx, ny, steps = int(g["nx"]), int(g["ny"]), int(g["steps"])
    u = orc.uniform_new(tau_default(), int(g["fluid_ty"]), nx * ny)
    s = orc.OracleSim(nx, ny, g["info"], u)
    for off, cell in zip(g["post_offsets"], g["post_cells"]):
        s.write_lattice_info(int(off), np.array([cell], W.LATTICE_INFO_DTYPE))
    if "particles" in g.files:
        num = tuple(int(v) for v in g["particle_num"])
        canvas_size = (nx * 2, ny * 2)
        from simuverse_b200.d2q9_node import SettingObj

        pu = SettingObj().particles_uniform_data
        pu.num[:] = list(num)
        parts = orc.init_trajectory_particles(canvas_size[0], canvas_size[1], num[0], num[1], pu.life_time, 0x5EED)
        assert parts.tobytes() == g["particles_init"].tobytes()
        canvas = np.zeros(canvas_size[0] * canvas_size[1], W.PIXEL_DTYPE)
        field = orc.field_uniform_new(nx, ny, 2, *canvas_size)
        for _ in range(steps):
            s.step(1)
            s.particle_update(field, pu, parts, canvas)
        assert parts.tobytes() == g["particles"].tobytes(), "particle trajectories"
        assert canvas.tobytes() == g["canvas"].tobytes(), "canvas"
    elif "midrun" in g.files:
        after, x0, x1, y0, y1 = (int(v) for v in g["midrun"])
        s.step(after)
        patch = np.zeros(x1 - x0, W.LATTICE_INFO_DTYPE)
        patch[:] = (W.OBSTACLE, -1, 0.0, 0.0)
        for y in range(y0, y1):
            s.write_lattice_info((y * nx + x0) * 16, patch)
        s.step(steps - after)
    else:
        s.step(steps)
    assert s.swap == int(g["swap"])
    # (the serial oracle visits cells in the same order as the WGSL harness, so even the slots that are
    # racy on a GPU agree here)
    assert_bits_equal(s.distributions(s.swap), g["buf_cur"], "current buffer")
    assert_bits_equal(s.distributions(1 - s.swap), g["buf_prev"], "previous buffer")
    np.testing.assert_array_equal(s.macro_f16, g["macro_f16"].reshape(-1))
    assert s.info.tobytes() == g["info_after"].tobytes()


def test_oracle_matches_reference_wgsl_live(orc):
    """When the reference tree is present (build container), transpile and run its shaders right here."""
    from wgsl_ref import harness as H

    if not H.available():
        pytest.skip("/root/reference is not present on this box")
    nx, ny = 14, 11
    info = orc.init_lattice_material(nx, ny, W.CUSTOM)
    g = info.reshape(ny, nx)
    g[5, 6] = (W.EXTERNAL_FORCE, -1, 0.05, -0.02)
    g["material"][3:5, 9:11] = W.OBSTACLE
    for fluid_ty in (0, 1):
        u = orc.uniform_new(tau_default(), fluid_ty, nx * ny)
        w = H.WgslLbm(nx, ny, info, u)
        s = orc.OracleSim(nx, ny, info, u)
        cell = np.array([(W.EXTERNAL_FORCE, 2, 0.03, 0.04)], W.LATTICE_INFO_DTYPE)
        w.info[3 * nx + 3] = cell[0]
        s.write_lattice_info((3 * nx + 3) * 16, cell)
        for _ in range(4):
            w.step(1)
            s.step(1)
            for b in (0, 1):
                assert_bits_equal(w.buf[b], s.buf[b], f"fluid_ty {fluid_ty} buf{b}")
            np.testing.assert_array_equal(w.macro.view(np.uint16).reshape(-1), s.macro_f16)
            assert w.info.tobytes() == s.info.tobytes()


def test_oracle_matches_executed_reference_wgsl_default_config(orc):
    """BASELINE configs[0] — the reference's default 600x375 lattice, its Poiseuille preset with the three
    R=28 discs and its 127x80 tracer particles — two frames of FluidSimulator::compute executed from the
    reference's WGSL source (tests/golden/make_wgsl_golden_default.py); digests of every buffer."""
    g = np.load(WGSL_DEFAULT)
    nx, ny, frames = int(g["nx"]), int(g["ny"]), int(g["frames"])
    canvas_size = (1200, 750)
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    s = orc.OracleSim(nx, ny, info, orc.uniform_new(tau_default(), 0, nx * ny), threads=4)
    from simuverse_b200.d2q9_node import SettingObj

    pu = SettingObj().particles_uniform_data
    pu.num[:] = [127, 80]
    parts = orc.init_trajectory_particles(canvas_size[0], canvas_size[1], 127, 80, pu.life_time, 0x5EED)
    canvas = np.zeros(canvas_size[0] * canvas_size[1], W.PIXEL_DTYPE)
    field = orc.field_uniform_new(nx, ny, 2, *canvas_size)
    for _ in range(2 * frames):
        s.step(1)
        s.particle_update(field, pu, parts, canvas)
    assert s.swap == int(g["swap"])
    assert sha(s.distributions(s.swap)) == str(g["sha_cur"]) and sha(s.distributions(1 - s.swap)) == str(g["sha_prev"])
    assert sha(s.macro_f16) == str(g["sha_macro"]) and sha(s.info) == str(g["sha_info"])
    assert sha(parts) == str(g["sha_particles"]) and sha(canvas) == str(g["sha_canvas"])
    rows = list(g["rows"])
    assert_bits_equal(s.distributions(s.swap)[:, rows, :], g["cur_rows"], "rows of the current buffer")
    assert abs(s.total_mass() - float(g["total_mass"])) < 1e-6


def test_curl_pass_matches_executed_reference_wgsl(orc):
    """oracle.curl_update against the texture produced by executing lbm/curl_update.wgsl (tests/golden/
    make_wgsl_golden_curl.py), plus the clamp quirk: the right / bottom taps of the last column / row are out of bounds
    and read zeros."""
    from helpers import WGSL_CURL

    g = np.load(WGSL_CURL)
    nx, ny = int(g["nx"]), int(g["ny"])
    got = orc.curl_update(nx, ny, g["macro_f16"])
    np.testing.assert_array_equal(got, g["curl_f16"])
    assert (got[..., 1:] == 0).all()
    # a field with u.y = 1 everywhere: interior curl = 0 -> 0.5; last column: right tap reads 0 -> curl = -1 -> -3.0;
    # column 0: left tap clamps onto itself -> 0 -> 0.5
    tex = np.zeros((5, 7, 4), np.float16)
    tex[..., 1] = 1.0
    c = orc.curl_update(7, 5, tex.view(np.uint16)).view(np.float16)[..., 0]
    assert (c[:, :6] == 0.5).all() and (c[:, 6] == -3.0).all()


def test_present_pass_matches_executed_reference_wgsl(orc):
    """oracle.present against the fragment outputs produced by executing lbm/present.wgsl per pixel (tests/golden/
    make_wgsl_golden_present.py) on two target sizes, bit for bit; plus hand-checked cases of the hsv2rgb / filter
    arithmetic, and a live re-run of a few rows when the reference tree is present."""
    from helpers import WGSL_CURL, WGSL_PRESENT

    g, c = np.load(WGSL_PRESENT), np.load(WGSL_CURL)
    nx, ny = int(g["nx"]), int(g["ny"])
    for tag in "ab":
        W, H = (int(v) for v in g[f"canvas_{tag}"])
        field = orc.field_uniform_new(nx, ny, 2, W, H)
        got = orc.present(field, c["macro_f16"], c["curl_f16"])
        assert_bits_equal(got, g[f"rgba_{tag}"], f"present target {W}x{H}")
        # a window of rows equals the same rows of the whole target
        assert_bits_equal(orc.present(field, c["macro_f16"], c["curl_f16"], 7, 5), g[f"rgba_{tag}"][7:12], "row window")
    # uniform textures: every filter weight sums to the texel value (weights (1-f)(1-g)+... are exact for f = g = 0.5 at
    # an integer ratio of 2), curl.x = 0.5 -> hue 0.5: p = |fract(0.5 + K) * 6 - 3| = (0, 1, 1) + rounding of 2/3, 1/3
    tex = np.zeros((4, 4, 4), np.float16)
    tex[..., 2] = 1.0  # rho = 1, u = 0
    curl = np.zeros((4, 4, 4), np.float16)
    curl[..., 0] = 0.5
    field = orc.field_uniform_new(4, 4, 2, 8, 8)
    out = orc.present(field, tex.view(np.uint16), curl.view(np.uint16))
    f = np.float32
    v = f(0.6) + f(1.0) * f(0.33)
    s = f(0.6)
    want_r = v * (f(1.0) * (f(1.0) - s) + f(0.0) * s)   # c = clamp(|fract(1.5) * 6 - 3| - 1) = clamp(-1) = 0
    assert (out[..., 3] == 1.0).all() and (out[..., 0] == want_r).all()
    assert np.all(out[..., 1] > out[..., 0]) and np.all(out[..., 2] > out[..., 0])  # cyan: g, b at full value
    from wgsl_ref import harness as H_

    if H_.available():
        from simuverse_b200.d2q9_node import lbm_uniform_new

        ch = np.load(WGSL_PRESENT.replace("wgsl_present_64x48", "wgsl_channel100_64x48_s100"))
        W, H = (int(v) for v in g["canvas_b"])
        sim = H_.WgslLbm(nx, ny, ch["info"], lbm_uniform_new(0.56, 0, nx * ny), canvas=(W, H))
        sim.macro[...] = c["macro_f16"].view(np.float16).reshape(ny, nx, 4)
        rows = [0, 50, H - 1]
        live = sim.present(c["curl_f16"].view(np.float16).reshape(ny, nx, 4), rows=rows)
        assert_bits_equal(live[rows], g["rgba_b"][rows], "live re-run of lbm/present.wgsl")
