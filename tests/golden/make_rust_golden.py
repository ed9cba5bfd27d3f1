"""Golden vectors of the reference's Rust HOST helpers on the LBM path, produced by EXECUTING their source text
(tests/rust_ref: a Rust-subset transpiler, f32 = numpy.float32, std math = libm) — the image has no Rust toolchain:

    init_lattice_material (fluid/lattice.rs:26-98)          masks of every preset at several sizes
    LbmUniform::new (fluid/mod.rs:32-55)                    the 304 uniform bytes
    FluidSimulator::update_uniforms (fluid_simulator.rs:175-193)   viscosity -> tau -> uniform bytes
    on_click -> add_obstacle (fluid_simulator.rs:137-152, d2q9_node.rs:215-245)   guard verdicts, write offsets, patches, mirror
    touch_begin / touch_move -> add_external_force (fluid_simulator.rs:154-173, d2q9_node.rs:263-300)   a drag: pre_pos
                                                            bookkeeping and every 16-byte write
    get_particles_data (lib.rs:247-273)                     tracer grid extent + workgroup count
    fullscreen_factor + the FieldUniform literal (util/matrix_helper.rs:19-41, d2q9_node.rs:61-76)   the 48 uniform bytes

Needs /root/reference (build container only); the output is committed.
Run from the repo root:  python tests/golden/make_rust_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from rust_ref import harness as H  # noqa: E402

POISEUILLE, LID, CUSTOM, BASIC = 4, 5, 6, 0
MASKS = [(600, 375, POISEUILLE), (96, 64, POISEUILLE), (131, 77, POISEUILLE), (257, 130, POISEUILLE), (48, 40, LID),
         (131, 77, LID), (40, 36, CUSTOM), (37, 19, BASIC)]
UNIFORMS = [(0.56, 0, 225000), (0.8, 1, 225000), (1.7, 0, 16777216), (0.5000001, 1, 9)]
VISCOSITIES = [(0.02, POISEUILLE), (0.1, LID), (0.0, CUSTOM), (0.3333, POISEUILLE)]
CLICKS = [(0.0, 10.0), (-1.0, 5.0), (55.9, 300.0), (56.0, 56.0), (600.0, 400.0), (610.0, 380.0), (1139.0, 689.0),
          (1140.0, 400.0), (700.5, 690.0), (80.0, 600.0)]
# a drag: touch_begin, then moves; (0, y) / negative coordinates reset pre_pos, jumps > 300 px only move it
DRAG = [(400.0, 300.0), (380.0, 310.0), (371.5, 333.25), (371.5, 333.25), (700.0, 333.0), (705.0, 330.0), (-3.0, 10.0),
        (10.0, 10.0), (12.0, 9.0), (3.0, 3.0), (1.0, 2.0), (1198.0, 700.0), (1100.0, 745.0), (1190.0, 748.0),
        (500.0, 0.0), (640.25, 100.75), (655.5, 118.0)]


def main():
    if not H.available():
        raise SystemExit("needs the reference tree at /root/reference")
    out = {"masks": np.array(MASKS, np.int32), "uniform_args": np.array(UNIFORMS, np.float64),
           "viscosities": np.array(VISCOSITIES, np.float64), "clicks": np.array(CLICKS, np.float32),
           "drag": np.array(DRAG, np.float32), "lattice": np.array([600, 375, 2], np.int32)}
    for k, (nx, ny, ty) in enumerate(MASKS):
        out[f"mask_{k}"] = H.init_lattice_material(nx, ny, ty)
        print("mask", nx, ny, ty, np.bincount(out[f"mask_{k}"]["material"]))
    out["uniform_bytes"] = np.frombuffer(b"".join(H.uniform_new(t, ty, n) for t, ty, n in UNIFORMS), np.uint8)
    nx, ny, lps = 600, 375, 2
    sim = H.Simulator(nx, ny, lps, POISEUILLE, out["mask_0"])
    for v, ty in VISCOSITIES:
        sim.update_uniforms(v, ty)
    out["update_uniform_bytes"] = np.frombuffer(b"".join(w[2] for w in sim.writes), np.uint8)
    del sim.writes[:]
    # clicks
    wrote, offs, patches = [], [], []
    for pos in CLICKS:
        n0 = len(sim.writes)
        sim.on_click(*pos)
        wrote.append(len(sim.writes) > n0)
        if wrote[-1]:
            assert sim.writes[-1][0] == "info_buf"
            offs.append(sim.writes[-1][1])
            patches.append(np.frombuffer(sim.writes[-1][2], H.LATTICE_INFO_DTYPE))
    out["click_wrote"] = np.array(wrote)
    out["click_offsets"] = np.array(offs, np.uint64)
    out["click_patches"] = np.stack(patches)
    out["mirror_after_clicks"] = H.info_to_array(sim.fluid_compute_node.lattice_info_data)
    print("clicks:", wrote, offs)
    # drag
    del sim.writes[:]
    sim.touch_begin()
    counts, pre = [], []
    for pos in DRAG:
        n0 = len(sim.writes)
        sim.touch_move(*pos)
        counts.append(len(sim.writes) - n0)
        pre.append((sim.pre_pos.x, sim.pre_pos.y))
    out["drag_write_counts"] = np.array(counts, np.int32)
    out["drag_pre_pos"] = np.array(pre, np.float32)
    out["drag_offsets"] = np.array([w[1] for w in sim.writes], np.uint64)
    out["drag_cells"] = np.frombuffer(b"".join(w[2] for w in sim.writes), H.LATTICE_INFO_DTYPE)
    print("drag writes per move:", counts)
    # FieldUniform as D2Q9Node::new builds it (d2q9_node.rs:38-76 + util/matrix_helper.rs:19-41)
    fields = [(1200, 750, 2), (750, 1200, 2), (2400, 1500, 4), (800, 800, 2), (333, 777, 3), (32768, 32768, 2)]
    out["field_args"] = np.array(fields, np.int32)
    out["field_bytes"] = np.frombuffer(b"".join(H.field_uniform(*f) for f in fields), np.uint8)
    # tracer grid (lib.rs:247-264)
    grids = [(1200, 750, 10000), (1200, 750, 205000), (2400, 1500, 40000), (800, 800, 1000), (333, 777, 5000),
             (16384, 16384, 1000000)]
    out["grid_args"] = np.array(grids, np.int32)
    out["grid_extent"] = np.array([H.particle_grid(*g)[0] for g in grids], np.int32)
    out["grid_workgroups"] = np.array([H.particle_grid(*g)[1] for g in grids], np.int32)
    print("grids:", out["grid_extent"].tolist())
    path = os.path.join(HERE, "rust_host_helpers.npz")
    # BASELINE configs[1]'s mask: 16.7 M cells through the interpreter take about a quarter of an hour, so only its CRC32
    # (of the i32-LE material array, the known answer of SURVEY 8c) and histogram are kept; `--crc4096` recomputes them
    if "--crc4096" in sys.argv:
        import zlib

        m = H.init_lattice_material(4096, 4096, POISEUILLE)
        assert (m["block_iter"] == -1).all() and (m["vy"] == 0).all()
        out["mask4096_crc"] = np.uint32(zlib.crc32(m["material"].astype("<i4").tobytes()))
        out["mask4096_hist"] = np.bincount(m["material"], minlength=8)
        out["mask4096_vx_sum"] = np.float64(m["vx"].astype(np.float64).sum())
        print("4096^2: crc %08x" % int(out["mask4096_crc"]), out["mask4096_hist"].tolist())
    elif os.path.exists(path):
        old = np.load(path)
        for k in ("mask4096_crc", "mask4096_hist", "mask4096_vx_sum"):
            if k in old.files:
                out[k] = old[k]
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
