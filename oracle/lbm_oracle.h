/*
 * lbm_oracle.h — CPU restatement of the reference's D2Q9 LBM path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and only as the checker / reported CPU baseline.
 *
 * HOW IT IS PINNED: the reference (Rust + WGSL on wgpu) has no tests, golden vectors or fixtures
 * for this path and cannot be built or run through wgpu here (no cargo/rustc,
 * no Vulkan ICD).  What CAN be executed is the reference's own shader text: tests/wgsl_ref
 * transpiles the unmodified assets/wgsl/lbm/{init,collide_stream,boundary,particle_update}.wgsl
 * (after the reference's #include expansion) to Python and evaluates them with IEEE f32 scalars;
 * tests/golden/make_wgsl_golden.py commits the resulting vectors (tests/golden/wgsl_*.npz) and
 * tests/test_oracle.py requires this oracle to reproduce them bit for bit (distributions, macro
 * texture, LatticeInfo, particles, canvas).  The Rust HOST helpers (LbmUniform::new,
 * init_lattice_material, add_obstacle, add_external_force, the click / drag guards, update_uniforms,
 * the tracer grid) are pinned the same way: tests/rust_ref executes their bodies cut out of the
 * reference's .rs files, tests/golden/make_rust_golden.py commits masks, uniform bytes, a click
 * sequence and a drag, and tests/test_rust_ref.py requires this oracle to reproduce every byte.
 * Not pinned: the tracer seeding (the reference draws from an unseeded rand::rng()).
 *
 * This file restates the reference's in-tree arithmetic — the WGSL shaders and the Rust host
 * functions cited per function — in plain C, f32, round-to-nearest, no FMA contraction
 * (-ffp-contract=off), evaluated in WGSL source order.  It is additionally cross-checked against
 * an independent numpy restatement (tests/np_restatement.py).
 */
#ifndef LBM_ORACLE_H
#define LBM_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/lbm_wire.h"

#ifdef __cplusplus
extern "C" {
#endif

/* f32 <-> f16 (IEEE binary16, round-to-nearest-even), what an rgba16float store/load does. */
uint16_t orc_f32_to_f16(float v);
float    orc_f16_to_f32(uint16_t h);

/* OpenMP threads used by the lattice passes (1 = serial; deterministic either way for
 * masks installed before init, see orc_boundary). Returns the value in effect. */
int orc_set_num_threads(int n);
int orc_get_max_threads(void);

/* fluid/mod.rs:31-55  LbmUniform::new */
void orc_lbm_uniform_new(float tau, int32_t fluid_ty, int32_t soa_offset, LbmUniform *out);

/* d2q9_node.rs:50 / fluid_simulator.rs:177  tau = 3*viscosity + 0.5 (f32) */
float orc_tau_from_viscosity(float viscosity);

/* fluid/lattice.rs:26-98  init_lattice_material (depth 1); ty is a FieldAnimationType */
void orc_init_lattice_material(int32_t nx, int32_t ny, int32_t ty, LatticeInfo *out);

/* Poiseuille frame whose Bulk cells become Obstacle iff
 * splitmix64(seed ^ (y<<32 | x)) top-24-bits / 2^24 < solid_fraction. Not a reference
 * function; restated here so the library's generator can be checked against it. */
void orc_init_porous_material(int32_t nx, int32_t ny, uint64_t seed, float solid_fraction,
                              LatticeInfo *out);

/* assets/wgsl/lbm/init.wgsl:19-63.  macro_f16: 4 halfs per cell (RGBA16F) or NULL. */
void orc_init(const LbmUniform *u, int32_t nx, int32_t ny, float *buf0, float *buf1,
              LatticeInfo *info, uint16_t *macro_f16);

/* assets/wgsl/lbm/collide_stream.wgsl:25-88 (+ layout_and_fn.wgsl:38-51, d2q9_fn.wgsl).
 * rd = collide_cell (read-only), wr = stream_cell.  macro_f16 (4 halfs/cell) and
 * macro_f32 (4 floats/cell: u.x,u.y,rho,1 before the f16 store) may each be NULL. */
void orc_collide_stream(const LbmUniform *u, int32_t nx, int32_t ny, const float *rd, float *wr,
                        LatticeInfo *info, uint16_t *macro_f16, float *macro_f32);

/* assets/wgsl/lbm/boundary.wgsl:3-35.  Row-major serial order when 1 thread. */
void orc_boundary(const LbmUniform *u, int32_t nx, int32_t ny, float *wr, const LatticeInfo *info);

/* d2q9_node.rs:302-312  compute_by_pass: collide_stream then boundary on the same bind group. */
void orc_step(const LbmUniform *u, int32_t nx, int32_t ny, const float *rd, float *wr,
              LatticeInfo *info, uint16_t *macro_f16, float *macro_f32);

/* n steps alternating buf0->buf1, buf1->buf0 starting with swap index `first_swap`
 * (fluid_simulator.rs:223-231 without the particle passes). Returns the swap index
 * the next step would use. */
int orc_step_n(const LbmUniform *u, int32_t nx, int32_t ny, float *buf0, float *buf1,
               LatticeInfo *info, uint16_t *macro_f16, float *macro_f32, int first_swap, int n);

/* assets/wgsl/lbm/particle_update.wgsl:55-88 + func/bilinear_interpolate_3f.wgsl.
 * Particles are visited in index order, so canvas collisions resolve to the highest index. */
void orc_particle_update(const LbmUniform *u, const FieldUniform *field, const ParticleUniform *pu,
                         TrajectoryParticle *particles, Pixel *canvas, const uint16_t *macro_f16);

/* assets/wgsl/present.wgsl:19-22,43-49: the in-place part of the canvas present pass (what
 * FluidSimulator::draw_by_rpass runs, fluid_simulator.rs:247): every pixel with alpha > 0.001 fades,
 * alpha *= fade_out_factor when alpha >= 0.2, else alpha *= 0.5.  (The fragment colour is render-only.) */
void orc_canvas_fade(const FieldUniform *field, const ParticleUniform *pu, Pixel *canvas);

/* d2q9_node.rs:215-245 add_obstacle. Mutates the host mirror; writes the patch the
 * reference uploads (rows [y-28, y+28) x nx) to `patch` (capacity 56*nx) and its byte
 * offset into the info buffer. Returns the number of LatticeInfo elements in the patch. */
size_t orc_add_obstacle(int32_t nx, int32_t ny, LatticeInfo *mirror, uint32_t x, uint32_t y,
                        LatticeInfo *patch, uint64_t *byte_offset);

/* fluid_simulator.rs:137-152 on_click guard: returns 1 and (x,y) lattice coords when the
 * click is accepted. */
int orc_on_click_guard(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, float px, float py,
                       uint32_t *x, uint32_t *y);

/* d2q9_node.rs:263-300 add_external_force. Emits up to `cap` (byte_offset, LatticeInfo)
 * single-cell writes in order; returns how many. */
size_t orc_add_external_force(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, float pos_x,
                              float pos_y, float pre_x, float pre_y, uint64_t *byte_offsets,
                              LatticeInfo *cells, size_t cap);

/* d2q9_node.rs:65-76 FieldUniform for the LBM node (proj_ratio / ndc_pixel from
 * util/matrix_helper.rs fullscreen_factor are render-only and left 0 here). */
void orc_field_uniform_new(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, int32_t canvas_w,
                           int32_t canvas_h, FieldUniform *out);

/* lib.rs:247-273 particle grid extent for (canvas, count): num.x, num.y */
void orc_particle_grid(uint32_t canvas_w, uint32_t canvas_h, int32_t count, int32_t *num_x,
                       int32_t *num_y);

/* lib.rs:275-316 init_trajectory_particles with a seeded splitmix64 stream in place of the
 * reference's unseeded rand::rng(). Index order is the reference's push
 * order: x outer, y inner. Writes num_x*num_y particles. */
void orc_init_trajectory_particles(uint32_t canvas_w, uint32_t canvas_h, int32_t num_x,
                                   int32_t num_y, float life_time, uint64_t seed,
                                   TrajectoryParticle *out);

/* assets/wgsl/lbm/curl_update.wgsl:12-33 — the derived-field pass FluidSimulator builds but never dispatches
 * (fluid_simulator.rs:226,230).  Per cell (no material test): curl = fb(right).y - fb(left).y + fb(top).x -
 * fb(bottom).x on the RGBA16F macro texture, left/top clamped to 0 and right/bottom to lattice_size — ONE PAST the
 * last texel (:21,24), where the load is out of bounds: zeros under wgpu (naga ReadZeroSkipWrite).  Stores
 * (curl * 3.5 + 0.5, 0, 0, 0) as f16 texels. */
void orc_curl_update(int32_t nx, int32_t ny, const uint16_t *macro_f16, uint16_t *curl_f16);

/* assets/wgsl/lbm/present.wgsl:21-46 — the colour present of the field (`render_node`, fluid_simulator.rs:69-87; built by
 * the reference, its draw call commented out at :243-244).  Fragment outputs (r, g, b, a) as f32 for canvas rows
 * [row0, row0 + rows) of a canvas_size[0] x canvas_size[1] target: hsv2rgb(curl.x, 0.6 + speed * 1.4, 0.6 + rho * 0.33),
 * alpha = rho, with macro / curl sampled bilinearly (ClampToEdge) at uv = pixel centre / canvas_size. */
void orc_present(const FieldUniform *field, const uint16_t *macro_f16, const uint16_t *curl_f16, int32_t row0,
                 int32_t rows, float *out_rgba);

/* f64 sum of one distribution buffer (all 9 planes) — the "total mass" diagnostic. */
double orc_total_mass(int32_t nx, int32_t ny, const float *buf);

#ifdef __cplusplus
}
#endif
#endif
