/*
 * lbm_b200.h — C ABI of the B200-native D2Q9 lattice-Boltzmann step.
 *
 * This is the drop-in boundary for the one hot path of jinleili/simuverse: what a
 * Rust `-sys` crate (rust/lbm-b200-sys), a ctypes binding (simuverse_b200/_capi.py) or
 * any other FFI binds in place of the wgpu objects owned by the reference's
 * `D2Q9Node` (simuverse/src/fluid/d2q9_node.rs:13-27).  Plain pointers and sizes
 * only; no torch / CUDA types in any signature.  Every function returns an
 * LbmStatus (0 = OK) and never unwinds across the boundary (the reference panics:
 * util/shader.rs:81,123).
 *
 * Correspondence with the reference (file:line relative to the reference tree):
 *
 *   lbm_create                 D2Q9Node::new buffer/texture creation     d2q9_node.rs:31-209
 *   lbm_write_uniform          queue.write_buffer(lbm_uniform_buf)       fluid_simulator.rs:188-192
 *   lbm_write_field_uniform    create_uniform_buffer(field_uniform_data) d2q9_node.rs:65-83
 *   lbm_write_lattice_info     queue.write_buffer(info_buf, off, bytes)  d2q9_node.rs:244,250-254,298
 *   lbm_reset                  reset_node.compute (init.wgsl)            d2q9_node.rs:211-213, init.wgsl:19-63
 *   lbm_step                   compute_by_pass(cpass, swap_index)        d2q9_node.rs:302-312
 *                              = collide_stream.wgsl:25-88 then boundary.wgsl:3-35, fused
 *   lbm_step_n                 the frame loop's alternation 0,1,0,1,...  fluid_simulator.rs:223-231
 *   lbm_compute_frames         FluidSimulator::compute, n times          fluid_simulator.rs:217-232
 *   lbm_read_macro             macro_tex (RGBA16F) contents              d2q9_node.rs:91-104
 *   lbm_read_curl              _curl_cal_node + curl_tex contents        fluid_simulator.rs:36-71,
 *                              = curl_update.wgsl:12-33
 *   lbm_read_present           render_node (colour present of the field) fluid_simulator.rs:69-87
 *                              = lbm/present.wgsl:21-46
 *   lbm_particles_update       particle_update_node.compute_by_pass      fluid_simulator.rs:225,229
 *                              = particle_update.wgsl:55-88
 *   lbm_read_distributions /   (no reference equivalent: can_read_back=false,
 *   lbm_write_distributions     d2q9_node.rs:112-117) — parity tests and checkpoint/restore
 *
 * Host-side (CPU, no GPU needed) mirrors of the reference's Rust helpers live at the
 * bottom: lbm_uniform_new, lbm_init_lattice_material, lbm_obstacle_patch, ...
 *
 * Threading: a handle is not thread-safe; one caller thread, like the reference's
 * single-threaded app (app_handler.rs:34).  All calls on a handle are stream-ordered;
 * reads synchronise before returning.
 *
 * Multi-GPU: one handle per GPU (one process per GPU under torch.distributed, or several
 * handles in one process).  Handle `rank` of `world` owns the y-slab
 * rows [ny*rank/world, ny*(rank+1)/world).  Neighbouring slabs are wired with
 * lbm_ipc_export / lbm_ipc_attach; after that the edge rows of every step read and write
 * the neighbours' memory directly over NVLink and steps are ordered by device-side flags.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stdint.h>
#include "lbm_wire.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_B200_ABI_VERSION 1

typedef struct LbmSim LbmSim; /* opaque */

typedef enum LbmStatus {
    LBM_OK = 0,
    LBM_ERR_INVALID_ARG = 1,
    LBM_ERR_CUDA = 2,          /* a CUDA runtime call failed; see lbm_last_error */
    LBM_ERR_NO_DEVICE = 3,     /* no usable CUDA device: there is NO CPU fallback */
    LBM_ERR_OUT_OF_MEMORY = 4,
    LBM_ERR_UNSUPPORTED = 5,   /* e.g. a uniform whose e-vectors are not D2Q9 */
    LBM_ERR_STATE = 6          /* call order violated (e.g. step before uniform upload) */
} LbmStatus;

/* LbmDesc.flags */
#define LBM_FLAG_MACRO_EVERY_STEP 0x1u /* write the RGBA16F macro texture in every step, like
                                          collide_stream.wgsl:74 (needed by the tracer particles).  Two-update
                                          sweeps store it too (the texture between the two updates only when
                                          tracer particles will read it).
                                          Off: the field is produced on demand by lbm_read_macro. */
#define LBM_FLAG_KERNEL_GENERIC   0x2u /* force the one-thread-per-cell kernel (A/B testing) */
#define LBM_FLAG_AA               0x8u /* AA-pattern in-place streaming: ONE copy of the distributions instead of
                                          the A/B ping-pong pair (half the memory, same traffic).  Single slab
                                          only; lbm_read_distributions / lbm_total_mass serve the current state
                                          (which == lbm_swap_index), canonicalised to the reference layout; the
                                          previous buffer and the on-demand macro field do not exist. */
#define LBM_FLAG_NO_GRAPH         0x4u /* launch every kernel individually instead of replaying CUDA
                                          graphs of 16 steps / one frame (A/B testing) */

#define LBM_FLAG_NO_FUSE          0x10u /* never run two updates as one sweep (lbm_fused.cuh); by default lbm_step_n and
                                          lbm_compute_frames do so whenever nothing has to happen between the two
                                          updates (no tracer particles / per-update macro texture, no force cell
                                          counting down).  Results are identical either way; the buffer that is not
                                          current is recomputed on demand when it is read (A/B testing) */

/* lbm_read_macro formats */
#define LBM_MACRO_F32_PLANES 0 /* 3 planes (u.x, u.y, rho) of rows*nx f32, before the f16 store */
#define LBM_MACRO_RGBA16F    1 /* rows*nx texels of 4 halfs (u.x, u.y, rho, 1), the reference texture */

/* lbm_generate_lattice_info presets beyond FieldAnimationType */
#define LBM_PRESET_POROUS 100 /* SURVEY.md §8d config 5 */

typedef struct LbmDesc {
    uint32_t struct_size;        /* = sizeof(LbmDesc) */
    int32_t  nx, ny;             /* global lattice (D2Q9Node::lattice, d2q9_node.rs:39-43) */
    int32_t  lattice_pixel_size; /* d2q9_node.rs:38 */
    int32_t  canvas_w, canvas_h; /* particle canvas in pixels; 0 = nx*lattice_pixel_size etc. */
    int32_t  device;             /* CUDA device ordinal; -1 = current device */
    int32_t  rank, world;        /* y-slab decomposition; world <= 1 means the whole lattice */
    uint32_t flags;              /* LBM_FLAG_* */
    int32_t  max_particles;      /* capacity of the particle buffer; 0 = no tracer particles */
} LbmDesc;

/* Opaque blob a slab publishes to its two y-neighbours (exchange it with any transport,
 * e.g. torch.distributed.all_gather). */
typedef struct LbmIpcBlob {
    uint8_t bytes[256];
} LbmIpcBlob;

/* ------------------------------------------------------------------ lifecycle */
int         lbm_abi_version(void);
int         lbm_device_count(void);                 /* 0 when no GPU / driver */
int         lbm_create(const LbmDesc *desc, LbmSim **out);
void        lbm_destroy(LbmSim *sim);
const char *lbm_last_error(const LbmSim *sim);      /* sim may be NULL: last lbm_create failure */
const char *lbm_status_string(int status);

/* ------------------------------------------------------------------ uploads */
int lbm_write_uniform(LbmSim *sim, const LbmUniform *u);
int lbm_write_field_uniform(LbmSim *sim, const FieldUniform *f);
/* byte_offset/nbytes address the GLOBAL info buffer (nx*ny*16 bytes, row-major); a slab keeps
 * the part that intersects its rows and one halo row each side, so every rank may be handed
 * the same call (on a multi-slab lattice it MUST be: the slabs derive from the bytes, each on its own, whether the
 * next updates may run as two-update sweeps — see lbm_scan_lattice_info_write).  Cost: O(nbytes) on the host, no
 * device round trip: the bytes are copied into a pinned staging ring (src may be reused at once) and uploaded
 * asynchronously, ordered before the next step. */
int lbm_write_lattice_info(LbmSim *sim, uint64_t byte_offset, const void *src, uint64_t nbytes);
/* Same content as lbm_init_lattice_material / lbm_init_porous_material, generated on the
 * device (no host array, no upload).  kind: FieldAnimationType or LBM_PRESET_POROUS. */
int lbm_generate_lattice_info(LbmSim *sim, int32_t kind, uint64_t seed, float solid_fraction);

/* ------------------------------------------------------------------ compute */
int lbm_reset(LbmSim *sim);                         /* init.wgsl; next step reads buffer 0 */
int lbm_step(LbmSim *sim, int32_t swap_index);      /* reads buffer swap_index, writes the other */
int lbm_step_n(LbmSim *sim, int32_t n);             /* n steps alternating from lbm_swap_index */
/* n_frames times FluidSimulator::compute (fluid_simulator.rs:217-232): step(0), particle update,
 * step(1), particle update (the particle updates only if the handle has tracer particles).  Unless something
 * mutates between the two updates (a force cell counting down) a frame is ONE launch of the two-update sweep kernel,
 * which also stores the macro texture of the second update — and of the first one into a second texture when tracer
 * particles will sample it — followed by the two particle passes: particles only read the field, so the results are
 * those of the reference's order, bit for bit. */
int lbm_compute_frames(LbmSim *sim, int32_t n_frames);
int lbm_swap_index(const LbmSim *sim);              /* buffer the next lbm_step_n step reads */
/* After a two-update sweep the buffer that is not current holds the state two updates back; the calls that need what
 * the reference has there (the state one update back: lbm_read_distributions / lbm_total_mass of that buffer,
 * the on-demand lbm_read_macro, lbm_write_*, lbm_step from it) recompute it first by themselves.  On a multi-slab
 * lattice that recomputation is an ordinary neighbour-synchronised update and therefore COLLECTIVE: call
 * lbm_refresh_previous on every slab, then synchronise all slabs (lbm_sync + host barrier), then read.  A no-op when
 * the last launch was not a sweep. */
int lbm_refresh_previous(LbmSim *sim);
int lbm_sync(LbmSim *sim);

/* ------------------------------------------------------------------ read-back / restore */
/* Rows owned by this handle: [*y0, *y0 + *rows). */
int lbm_slab_rows(const LbmSim *sim, int32_t *y0, int32_t *rows);
/* dst/src: 9 planes of rows*nx f32, plane-major like the reference buffer
 * (d2q9_fn.wgsl:13-17) restricted to the owned rows. which = 0 or 1. */
int lbm_read_distributions(LbmSim *sim, int32_t which, float *dst);
int lbm_write_distributions(LbmSim *sim, int32_t which, const float *src);
int lbm_read_macro(LbmSim *sim, int32_t format, void *dst);
/* Pipelined variant for per-frame consumers (handle created with LBM_FLAG_MACRO_EVERY_STEP): enqueues the
 * device-to-host copy of the RGBA16F texture on a separate stream and returns; later steps write a second
 * texture meanwhile.  dst (pinned host memory) is complete after lbm_sync. */
int lbm_read_macro_async(LbmSim *sim, void *dst);
/* The derived-field pass of the reference, lbm/curl_update.wgsl:12-33 (`_curl_cal_node`, fluid_simulator.rs:36-71;
 * built there but never dispatched, :226,230): curl of the velocity in the newest macro texture, stored like the
 * reference's `curl_tex` as RGBA16F texels (curl * 3.5 + 0.5, 0, 0, 0).  dst: rows*nx texels of 4 halfs.  The
 * shader's right / bottom taps are clamped to lattice_size — one past the last texel — where wgpu reads zeros; that is
 * reproduced.  Single-slab handles only (the reference is single-GPU and nothing consumes the texture). */
int lbm_read_curl(LbmSim *sim, void *dst);
/* The colour present of the field, lbm/present.wgsl:21-46 (`render_node`, fluid_simulator.rs:69-87; built by the
 * reference, its draw call commented out at :243-244): the fragment outputs (r, g, b, a) as f32 for rows
 * [row0, row0 + rows) of a FieldUniform.canvas_size target — hsv2rgb(curl.x, 0.6 + speed * 1.4, 0.6 + rho * 0.33), alpha
 * = rho, the newest macro texture and its curl sampled through `bilinear_sampler` (ClampToEdge, linear) at
 * uv = pixel centre / canvas_size.  The filter is WebGPU's formula in f32, one rounding per operation:
 * c = uv * size - 0.5, i = floor(c), f = c - i, taps clamped to the edge, (1-fx)(1-fy) t00 + fx(1-fy) t10 + (1-fx) fy t01
 * + fx fy t11 summed in that order (hardware samplers use fixed-point weights of unspecified width, so no GPU's output is
 * bit-comparable with another's; the tests pin this arithmetic on the executed shader text).  The surface-format
 * conversion of the render target is the caller's.  dst: rows * canvas_size[0] * 4 floats.
 * Single-slab handles only. */
int lbm_read_present(LbmSim *sim, int32_t row0, int32_t rows, float *dst);
/* Owned rows of the info buffer including device-side block_iter/material mutation
 * (collide_stream.wgsl:55-62). dst: rows*nx LatticeInfo. */
int lbm_read_lattice_info(LbmSim *sim, LatticeInfo *dst);
/* f64 sum over the owned rows of all 9 planes of buffer `which`. */
int lbm_total_mass(LbmSim *sim, int32_t which, double *out);

/* ------------------------------------------------------------------ tracer particles */
int lbm_write_particle_uniform(LbmSim *sim, const ParticleUniform *pu);
int lbm_particles_write(LbmSim *sim, const TrajectoryParticle *src, uint64_t count);
int lbm_particles_update(LbmSim *sim);              /* particle_update.wgsl:55-88 */
int lbm_particles_read(LbmSim *sim, TrajectoryParticle *dst, uint64_t count);
int lbm_canvas_clear(LbmSim *sim);
/* The in-place part of the canvas present pass run by FluidSimulator::draw_by_rpass
 * (fluid_simulator.rs:247, present.wgsl:19-22,43-49): alpha fade of every lit pixel. */
int lbm_canvas_fade(LbmSim *sim);
int lbm_canvas_read(LbmSim *sim, Pixel *dst);       /* canvas_w*canvas_h pixels */

/* ------------------------------------------------------------------ multi-GPU wiring */
int lbm_ipc_export(LbmSim *sim, LbmIpcBlob *out);
/* up = slab owning row y0-1 (rank-1 mod world), down = slab owning row y0+rows. */
int lbm_ipc_attach(LbmSim *sim, const LbmIpcBlob *up, const LbmIpcBlob *down);

/* ------------------------------------------------------------------ timing / introspection */
/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
uint64_t lbm_launch_count(const LbmSim *sim);
/* How many of those launches were two-update sweeps (k_frame2), i.e. advanced the lattice by two updates. */
uint64_t lbm_fused_sweep_count(const LbmSim *sim);
/* 1 if sweeps run the kernel instance with the inline masked path for cells next to solids (the default), 0 if
 * LBM_FUSE_MASKED=0 was in the environment at lbm_create (A/B runs, tests).  Both instances produce identical results. */
int lbm_sweep_uses_masked_path(const LbmSim *sim);
/* Multi-slab diagnostics (synchronises): nanoseconds the edge CTAs of this slab have spent waiting for their neighbour
 * slabs' progress flags since creation, summed over CTAs, and the number of such waits. */
int lbm_edge_wait_stats(LbmSim *sim, uint64_t *total_ns, uint64_t *n_waits);
/* Device time of the last lbm_step_n call in milliseconds, measured with CUDA events on the
 * handle's own stream (torch.cuda.Event only sees torch's current stream). */
int lbm_last_step_n_ms(LbmSim *sim, float *ms);
/* Raw stream handle (cudaStream_t) for callers that want to order their own work. */
void *lbm_stream(LbmSim *sim);

/* ------------------------------------------------------------------ host-side mirrors (CPU only) */
/* Row blocks of a two-update sweep in dispatch order (csrc/lbm_fused.cuh): (y0, y1) pairs into out[2*cap]; returns the
 * number of blocks (0 on bad arguments / too small cap), *n_edge = how many leading blocks hold the slab's first / last
 * rows.  No reference counterpart; exported so that the cutting rules are testable without a GPU. */
int32_t lbm_sweep_blocks(int32_t h, int32_t rows_per_block, int32_t *out, int32_t cap, int32_t *n_edge);
/* Same with the last tail_rows rows cut into blocks of tail_rows_per_block rows (dispatched last: short work items at the
 * end of a grid of few waves). */
int32_t lbm_sweep_blocks_tail(int32_t h, int32_t rows_per_block, int32_t tail_rows, int32_t tail_rows_per_block,
                              int32_t *out, int32_t cap, int32_t *n_edge);
/* The schedule verdict of a lbm_write_lattice_info call, as a pure function of its arguments (every rank of a multi-slab
 * lattice computes it from the same bytes): *armed = largest block_iter > 0 of a written inlet / force cell,
 * *border_solid = 1 if a solid is written within one cell of the outer ring.  Returns 0 (and the conservative answer)
 * for writes that are not LatticeInfo-aligned.  No reference counterpart; exported for CPU tests. */
int32_t lbm_scan_lattice_info_write(int32_t nx, int32_t ny, uint64_t byte_offset, const void *src, uint64_t nbytes,
                                    int32_t *armed, int32_t *border_solid);
/* fluid/mod.rs:31-55 LbmUniform::new */
void  lbm_uniform_new(float tau, int32_t fluid_ty, int32_t soa_offset, LbmUniform *out);
/* d2q9_node.rs:50, fluid_simulator.rs:177 */
float lbm_tau_from_viscosity(float viscosity);
/* d2q9_node.rs:65-76 (proj_ratio / ndc_pixel are render-only and left 0) */
void  lbm_field_uniform_new(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, int32_t canvas_w,
                            int32_t canvas_h, FieldUniform *out);
/* fluid/lattice.rs:26-98 */
int   lbm_init_lattice_material(int32_t nx, int32_t ny, int32_t ty, LatticeInfo *out);
int   lbm_init_porous_material(int32_t nx, int32_t ny, uint64_t seed, float solid_fraction,
                               LatticeInfo *out);
/* fluid_simulator.rs:137-152 on_click guard; returns 1 when accepted */
int   lbm_on_click_guard(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, float px, float py,
                         uint32_t *x, uint32_t *y);
/* d2q9_node.rs:215-245 add_obstacle: updates the CPU mirror, fills the 56-row patch and its
 * byte offset; returns the element count of the patch. */
uint64_t lbm_obstacle_patch(int32_t nx, int32_t ny, LatticeInfo *mirror, uint32_t x, uint32_t y,
                            LatticeInfo *patch, uint64_t *byte_offset);
/* d2q9_node.rs:263-300 add_external_force: the single-cell writes it issues, in order */
uint64_t lbm_external_force_cells(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, float pos_x,
                                  float pos_y, float pre_x, float pre_y, uint64_t *byte_offsets,
                                  LatticeInfo *cells, uint64_t cap);
/* lib.rs:247-264 */
void  lbm_particle_grid(uint32_t canvas_w, uint32_t canvas_h, int32_t count, int32_t *num_x,
                        int32_t *num_y);
/* lib.rs:275-316 with a seeded stream instead of rand::rng() */
void  lbm_init_trajectory_particles(uint32_t canvas_w, uint32_t canvas_h, int32_t num_x,
                                    int32_t num_y, float life_time, uint64_t seed,
                                    TrajectoryParticle *out);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
