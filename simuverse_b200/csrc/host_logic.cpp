// host_logic.cpp — CPU-side pieces of the drop-in: the reference's Rust host helpers for the
// LBM node restated in C++ (no GPU involved).  They produce the bytes the reference uploads
// with queue.write_buffer; the device path consumes them through lbm_write_*.
//
//   LbmUniform::new            simuverse/src/fluid/mod.rs:31-55
//   init_lattice_material      simuverse/src/fluid/lattice.rs:26-98
//   D2Q9Node::add_obstacle     simuverse/src/fluid/d2q9_node.rs:215-245
//   D2Q9Node::add_external_force                         d2q9_node.rs:263-300
//   FluidSimulator::on_click   simuverse/src/fluid/fluid_simulator.rs:137-152
//   get_particles_data / init_trajectory_particles       simuverse/src/lib.rs:247-316
//
// f32 throughout, one rounding per operation (built with -ffp-contract=off), so the masks are
// bit-identical to what the Rust code computes with glam::Vec2.
#include <vector>
#include <cmath>
#include <cstring>

#include "../../include/lbm_b200.h"

namespace {

constexpr float kObstacleRadius = 28.0f; // fluid/mod.rs:1

struct V2 {
    float x, y;
    V2 operator-(V2 o) const { return {x - o.x, y - o.y}; }
    float length() const { return std::sqrt(x * x + y * y); } // glam: sqrt(dot(self, self))
};

inline bool is_sd_sphere(V2 p, float r) { return p.length() <= r; } // fluid/mod.rs:57-59

inline LatticeInfo cell(int32_t material, float vx) { return LatticeInfo{material, -1, vx, 0.0f}; }

// Rust `as u32` on f32: saturating, NaN -> 0
inline uint32_t as_u32(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967296.0f) return 0xffffffffu;
    return static_cast<uint32_t>(v);
}
inline int32_t as_i32(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return -2147483647 - 1;
    return static_cast<int32_t>(v);
}

inline uint64_t mix64(uint64_t z) { // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

struct Stream { // seeded replacement for rand::rng()
    uint64_t s;
    float unit() {
        s += 0x9e3779b97f4a7c15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z ^= z >> 31;
        return static_cast<float>(z >> 40) / 16777216.0f;
    }
    float uniform(float lo, float hi) { return lo + (hi - lo) * unit(); }
};

// Poiseuille frame shared by the channel and porous presets (lattice.rs:65-76)
inline bool poiseuille_frame(int32_t x, int32_t y, int32_t nx, int32_t ny, LatticeInfo &out) {
    if (y == 0 || y == ny - 1) { out = cell(LATTICE_BOUNDARY, 0.0f); return true; }
    if (x == 0 || x == nx - 1) { out = cell(LATTICE_GHOST, 0.0f); return true; }
    if (x == 1) { out = cell(LATTICE_INLET, 0.12f); return true; }
    if (x == nx - 2) { out = cell(LATTICE_OUTLET, 0.0f); return true; }
    return false;
}

}  // namespace

extern "C" void lbm_uniform_new(float tau, int32_t fluid_ty, int32_t soa_offset, LbmUniform *out) {
    const float wc = 0.444444f, wa = 0.111111f, wd = 0.0277777f;
    const float mc = 0.6f, ma = 0.2222f, md = 0.1111f;
    const float table[9][4] = {
        {0.f, 0.f, wc, mc},  {1.f, 0.f, wa, ma},   {0.f, -1.f, wa, ma}, {-1.f, 0.f, wa, ma}, {0.f, 1.f, wa, ma},
        {1.f, -1.f, wd, md}, {-1.f, -1.f, wd, md}, {-1.f, 1.f, wd, md}, {1.f, 1.f, wd, md}};
    const int32_t opposite[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    out->tau = tau;
    out->omega = 1.0f / tau;
    out->fluid_ty = fluid_ty;
    out->soa_offset = soa_offset;
    std::memcpy(out->e_w_max, table, sizeof(table));
    for (int i = 0; i < 9; i++)
        for (int k = 0; k < 4; k++) out->inversed_direction[i][k] = opposite[i];
}

extern "C" float lbm_tau_from_viscosity(float viscosity) { return 3.0f * viscosity + 0.5f; }

extern "C" void lbm_field_uniform_new(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, int32_t canvas_w,
                                      int32_t canvas_h, FieldUniform *out) {
    FieldUniform f{};
    f.lattice_size[0] = nx;
    f.lattice_size[1] = ny;
    f.lattice_pixel_size[0] = f.lattice_pixel_size[1] = static_cast<float>(lattice_pixel_size);
    f.canvas_size[0] = canvas_w;
    f.canvas_size[1] = canvas_h;
    // proj_ratio / ndc_pixel (d2q9_node.rs:61-73 with util/matrix_helper.rs:27-37 fullscreen_factor): no LBM shader reads
    // them (they feed the field simulator's velocity code), but they are part of the 48 bytes the reference uploads
    const float vw = static_cast<float>(canvas_w), vh = static_cast<float>(canvas_h);
    float sx = 1.0f, sy = 1.0f;
    if (vh > vw) sy = vh / vw; else sx = vw / vh;
    f.proj_ratio[0] = sx;
    f.proj_ratio[1] = sy;
    f.ndc_pixel[0] = sx * 2.0f / vw;
    f.ndc_pixel[1] = sy * 2.0f / vh;
    f.speed_ty = 1;
    *out = f;
}

extern "C" int lbm_init_lattice_material(int32_t nx, int32_t ny, int32_t ty, LatticeInfo *out) {
    if (!out || nx < 1 || ny < 1) return LBM_ERR_INVALID_ARG;
    const float fnx = static_cast<float>(nx), fny = static_cast<float>(ny);
    const V2 discs[3] = {{fnx / 7.0f - kObstacleRadius, fny / 2.0f}, {fnx / 5.0f, fny / 4.0f}, {fnx / 5.0f, fny * 0.75f}};
    LatticeInfo *o = out;
    for (int32_t y = 0; y < ny; y++) {
        for (int32_t x = 0; x < nx; x++, o++) {
            *o = cell(LATTICE_BULK, 0.0f);
            switch (ty) {
                case FIELD_ANIMATION_POISEUILLE: {
                    if (poiseuille_frame(x, y, nx, ny, *o)) break;
                    const V2 p{static_cast<float>(x), static_cast<float>(y)};
                    if (is_sd_sphere(p - discs[0], kObstacleRadius) || is_sd_sphere(p - discs[1], kObstacleRadius) ||
                        is_sd_sphere(p - discs[2], kObstacleRadius))
                        o->material = LATTICE_OBSTACLE;
                    break;
                }
                case FIELD_ANIMATION_LID_DRIVEN_CAVITY:
                    if (x == 0 || x == nx - 1 || y == ny - 1) o->material = LATTICE_BOUNDARY;
                    else if (y == 0) o->material = LATTICE_GHOST;
                    else if (y == 1) *o = cell(LATTICE_EXTERNAL_FORCE, 0.13f);
                    break;
                case FIELD_ANIMATION_CUSTOM:
                    if (x == 0 || x == nx - 1 || y == 0 || y == ny - 1) o->material = LATTICE_BOUNDARY;
                    break;
                default: // the reference's `_ => {}`: all bulk
                    break;
            }
        }
    }
    return LBM_OK;
}

extern "C" int lbm_init_porous_material(int32_t nx, int32_t ny, uint64_t seed, float solid_fraction, LatticeInfo *out) {
    if (!out || nx < 1 || ny < 1) return LBM_ERR_INVALID_ARG;
    LatticeInfo *o = out;
    for (int32_t y = 0; y < ny; y++) {
        for (int32_t x = 0; x < nx; x++, o++) {
            *o = cell(LATTICE_BULK, 0.0f);
            if (poiseuille_frame(x, y, nx, ny, *o)) continue;
            const uint64_t key = (static_cast<uint64_t>(static_cast<uint32_t>(y)) << 32) | static_cast<uint32_t>(x);
            const float r = static_cast<float>(mix64(seed ^ key) >> 40) / 16777216.0f;
            if (r < solid_fraction) o->material = LATTICE_OBSTACLE;
        }
    }
    return LBM_OK;
}

extern "C" int lbm_on_click_guard(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, float px, float py, uint32_t *x,
                                  uint32_t *y) {
    if (px <= 0.0f || py <= 0.0f) return 0;
    const uint32_t cx = as_u32(px) / lattice_pixel_size, cy = as_u32(py) / lattice_pixel_size;
    const uint32_t half_size = static_cast<uint32_t>(kObstacleRadius);
    if (cx < half_size || cx >= static_cast<uint32_t>(nx) - (half_size + 2) || cy < half_size ||
        cy >= static_cast<uint32_t>(ny) - (half_size + 2))
        return 0;
    if (x) *x = cx;
    if (y) *y = cy;
    return 1;
}

extern "C" uint64_t lbm_obstacle_patch(int32_t nx, int32_t ny, LatticeInfo *mirror, uint32_t x, uint32_t y,
                                       LatticeInfo *patch, uint64_t *byte_offset) {
    (void)ny;
    const LatticeInfo obstacle = cell(LATTICE_OBSTACLE, 0.0f);
    const V2 center{static_cast<float>(x) + 0.5f, static_cast<float>(y) + 0.5f};
    const uint32_t r = static_cast<uint32_t>(kObstacleRadius);
    const uint32_t min_y = y - r, max_y = min_y + 2 * r;
    uint64_t n = 0;
    for (uint32_t yy = min_y; yy < max_y; yy++) {
        LatticeInfo *row = mirror + static_cast<size_t>(nx) * yy;
        for (uint32_t xx = 0; xx < static_cast<uint32_t>(nx); xx++) {
            const V2 p{static_cast<float>(xx) + 0.5f, static_cast<float>(yy) + 0.5f};
            if (is_sd_sphere(p - center, kObstacleRadius)) row[xx] = obstacle;
            patch[n++] = row[xx];
        }
    }
    if (byte_offset) *byte_offset = static_cast<uint64_t>(static_cast<uint32_t>(nx) * min_y) * sizeof(LatticeInfo);
    return n;
}

extern "C" uint64_t lbm_external_force_cells(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, float pos_x,
                                             float pos_y, float pre_x, float pre_y, uint64_t *byte_offsets,
                                             LatticeInfo *cells, uint64_t cap) {
    const V2 pos{pos_x, pos_y}, pre{pre_x, pre_y};
    const float dis = (pos - pre).length();
    float force = 0.1f * (dis / 20.0f);
    if (force > 0.12f) force = 0.12f;
    const float angle = std::atan2(pos.y - pre.y, pos.x - pre.x);
    const LatticeInfo forced{LATTICE_EXTERNAL_FORCE, 90, force * std::cos(angle), force * std::sin(angle)};
    // d2q9_node.rs:277 divides by lattice_pixel_size - 1.  With lattice_pixel_size == 1 (scale_factor <= 0.5) that is a
    // division by zero: c = +inf saturates to i32::MAX and the reference walks ~2^31 sample points.  Deliberate
    // deviation: a library entry point must not stall, so such a drag produces no force cells.
    if (lattice_pixel_size < 2) return 0;
    const float c = std::ceil(dis / static_cast<float>(lattice_pixel_size - 1));
    if (!std::isfinite(c)) return 0;
    const float step = dis / c;
    uint64_t n = 0;
    const int32_t count = as_i32(c);
    for (int32_t i = 0; i < count; i++) {
        const float d = step * static_cast<float>(i);
        const float qx = std::round(pre.x + d * std::cos(angle)), qy = std::round(pre.y + d * std::sin(angle));
        const uint32_t x = as_u32(qx) / lattice_pixel_size, y = as_u32(qy) / lattice_pixel_size;
        if (x < 1 || x >= static_cast<uint32_t>(nx) - 2 || y < 1 || y >= static_cast<uint32_t>(ny) - 2) continue;
        if (n < cap) {
            byte_offsets[n] = static_cast<uint64_t>(static_cast<uint32_t>(nx) * y + x) * sizeof(LatticeInfo);
            cells[n] = forced;
        }
        n++;
    }
    return n;
}

extern "C" void lbm_particle_grid(uint32_t canvas_w, uint32_t canvas_h, int32_t count, int32_t *num_x, int32_t *num_y) {
    const float ratio = static_cast<float>(canvas_w) / static_cast<float>(canvas_h);
    const float x = std::ceil(std::sqrt(static_cast<float>(count) * ratio));
    *num_x = static_cast<int32_t>(as_u32(x));
    *num_y = static_cast<int32_t>(as_u32(std::ceil(x * (1.0f / ratio))));
}

extern "C" void lbm_init_trajectory_particles(uint32_t canvas_w, uint32_t canvas_h, int32_t num_x, int32_t num_y,
                                              float life_time, uint64_t seed, TrajectoryParticle *out) {
    Stream rng{seed};
    const float step_x = static_cast<float>(canvas_w) / static_cast<float>(num_x - 1);
    const float step_y = static_cast<float>(canvas_h) / static_cast<float>(num_y - 1);
    const float life_hi = life_time <= 0.0f ? 1.0f : life_time;
    TrajectoryParticle *o = out;
    for (int32_t x = 0; x < num_x; x++) {
        const float pixel_x = step_x * static_cast<float>(x);
        for (int32_t y = 0; y < num_y; y++, o++) {
            const float jx = rng.uniform(-step_x, step_x);
            const float jy = rng.uniform(-step_y, step_y);
            o->pos[0] = pixel_x + jx;
            o->pos[1] = step_y * static_cast<float>(y) + jy;
            if (life_time <= 1.0f) {
                o->pos_initial[0] = rng.uniform(0.0f, step_x);
                o->pos_initial[1] = o->pos[1];
                o->life_time = 0.0f;
            } else {
                o->pos_initial[0] = o->pos[0];
                o->pos_initial[1] = o->pos[1];
                o->life_time = rng.uniform(0.0f, life_hi);
            }
            o->fade = 0.0f;
        }
    }
}

// ------------------------------------------------------------------ work items of a two-update sweep
// Cuts rows [0, h) into blocks of `rows_per_block` rows in DISPATCH order: the block with row 0, the block with row
// h-1, then the rest top to bottom (on a multi-slab lattice the first two are the ones that talk to the neighbour
// slabs, lbm_fused.cuh).  A 1-row remainder is merged into its predecessor, so that the last block always holds
// the slab's last two rows.  out: (y0, y1) pairs; returns the number of blocks, *n_edge = 1 or 2.
// lbm_sweep_blocks_tail: the last tail_rows rows are cut into shorter blocks (tail_rows_per_block).  They are the last
// CTAs dispatched: a grid of only a few waves then ends with short work items instead of a ragged last wave.
extern "C" int32_t lbm_sweep_blocks_tail(int32_t h, int32_t rows_per_block, int32_t tail_rows, int32_t tail_rows_per_block,
                                         int32_t *out, int32_t cap, int32_t *n_edge) {
    if (h < 1 || rows_per_block < 1 || !out) return 0;
    if (tail_rows < 0 || tail_rows >= h || tail_rows_per_block < 1) tail_rows = 0;
    std::vector<int32_t> lo_, hi_;
    const int32_t body = h - tail_rows;
    for (int32_t y = 0; y < body; y += rows_per_block) {
        lo_.push_back(y);
        hi_.push_back(y + rows_per_block < body ? y + rows_per_block : body);
    }
    for (int32_t y = body; y < h; y += tail_rows_per_block) {
        lo_.push_back(y);
        hi_.push_back(y + tail_rows_per_block < h ? y + tail_rows_per_block : h);
    }
    if (lo_.size() > 1 && hi_.back() - lo_.back() < 2) {
        lo_.pop_back();
        hi_.pop_back();
        hi_.back() = h;
    }
    const int32_t n = (int32_t)lo_.size();
    if (n > cap) return 0;
    int32_t k = 0;
    out[2 * k] = lo_.front(); out[2 * k + 1] = hi_.front(); k++;
    if (n > 1) { out[2 * k] = lo_.back(); out[2 * k + 1] = hi_.back(); k++; }
    for (int32_t i = 1; i + 1 < n; i++) { out[2 * k] = lo_[i]; out[2 * k + 1] = hi_[i]; k++; }
    if (n_edge) *n_edge = n > 1 ? 2 : 1;
    return n;
}

extern "C" int32_t lbm_sweep_blocks(int32_t h, int32_t rows_per_block, int32_t *out, int32_t cap, int32_t *n_edge) {
    return lbm_sweep_blocks_tail(h, rows_per_block, 0, 1, out, cap, n_edge);
}

// What the bytes of a lbm_write_lattice_info call mean for the step schedule (lbm_b200.cu).  Every rank of a multi-slab
// lattice is handed the same call and must reach the same verdict without communication, so it is a pure function of
// the arguments: *armed = the largest block_iter > 0 of an inlet / force cell among the written cells (single updates
// while it counts down, collide_stream.wgsl:55-62), *border_solid = 1 if a solid (material 2 or 4) is written within
// one cell of the outer ring — only there can a solid painted over live fluid leave a value that a ring cell keeps
// pulling (boundary.wgsl:19 never rewrites those slots).  Writes that are not LatticeInfo-aligned cannot be parsed:
// they report the conservative answer (border_solid = 1) and return 0.
extern "C" int32_t lbm_scan_lattice_info_write(int32_t nx, int32_t ny, uint64_t byte_offset, const void *src, uint64_t nbytes,
                                               int32_t *armed, int32_t *border_solid) {
    if (armed) *armed = 0;
    if (border_solid) *border_solid = 1;
    if (nx < 1 || ny < 1 || (!src && nbytes)) return 0;
    if (byte_offset % sizeof(LatticeInfo) != 0 || nbytes % sizeof(LatticeInfo) != 0) return 0;
    const LatticeInfo *cells = static_cast<const LatticeInfo *>(src);
    const uint64_t n = nbytes / sizeof(LatticeInfo);
    const uint64_t idx = byte_offset / sizeof(LatticeInfo);
    int32_t x = static_cast<int32_t>(idx % static_cast<uint64_t>(nx)), y = static_cast<int32_t>(idx / static_cast<uint64_t>(nx));
    int32_t a = 0, b = 0;
    for (uint64_t k = 0; k < n; k++) {
        const int32_t m = cells[k].material;
        if ((m == LATTICE_INLET || m == LATTICE_EXTERNAL_FORCE) && cells[k].block_iter > a) a = cells[k].block_iter;
        if ((m == LATTICE_BOUNDARY || m == LATTICE_OBSTACLE) && (x <= 1 || x >= nx - 2 || y <= 1 || y >= ny - 2)) b = 1;
        if (++x == nx) { x = 0; y++; }
    }
    if (armed) *armed = a;
    if (border_solid) *border_solid = b;
    return 1;
}
