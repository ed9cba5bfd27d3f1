/*
 * lbm_oracle.c — CPU restatement of the reference's D2Q9 LBM path (see lbm_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY — never linked into, loaded by, or called from the product.
 * Pinned to the reference's WGSL source executed by tests/wgsl_ref and to its Rust host helpers
 * executed by tests/rust_ref (see lbm_oracle.h).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math [-fopenmp]  (oracle/Makefile).
 * Every expression is written in the reference's source order so that, with FMA
 * contraction off, each f32 operation rounds exactly once in the same place as a
 * non-contracting WGSL implementation would.  Citations are reference file:line
 * relative to the reference tree.
 */
#include "lbm_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;

int orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n < 1) n = 1;
    g_threads = n;
#else
    (void)n;
    g_threads = 1;
#endif
    return g_threads;
}

int orc_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------- f16 <-> f32 */

uint16_t orc_f32_to_f16(float v) {
    uint32_t x;
    memcpy(&x, &v, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? (0x0200u | ((ax >> 13) & 0x03ffu)) : 0u));
    }
    if (ax >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x33000001u) { /* <= 2^-25: rounds to zero (2^-25 exactly ties to even = 0) */
        return (uint16_t)sign;
    }
    int32_t e = (int32_t)(ax >> 23) - 127;
    uint32_t m = (ax & 0x007fffffu) | 0x00800000u; /* 24-bit significand */
    uint32_t shift;
    uint32_t base;
    if (e < -14) { /* subnormal half */
        shift = (uint32_t)(13 + (-14 - e));
        base = 0;
    } else {
        shift = 13;
        base = (uint32_t)(e + 15) << 10;
        m &= 0x007fffffu;
    }
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q += 1; /* carries into the exponent correctly */
    return (uint16_t)(sign | (base + q));
}

float orc_f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) {
            x = sign;
        } else { /* subnormal: normalise */
            int s = 0;
            while (!(m & 0x400u)) { m <<= 1; s++; }
            m &= 0x3ffu;
            x = sign | ((uint32_t)(127 - 15 - s + 1) << 23) | (m << 13);
        }
    } else if (e == 31) {
        x = sign | 0x7f800000u | (m << 13);
    } else {
        x = sign | ((e + 127 - 15) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}

/* ---------------------------------------------------------------- host-side Rust pieces */

/* fluid/mod.rs:31-55 */
void orc_lbm_uniform_new(float tau, int32_t fluid_ty, int32_t soa_offset, LbmUniform *out) {
    static const float ewm[9][4] = {
        {0.0f, 0.0f, 0.444444f, 0.6f},      {1.0f, 0.0f, 0.111111f, 0.2222f},
        {0.0f, -1.0f, 0.111111f, 0.2222f},  {-1.0f, 0.0f, 0.111111f, 0.2222f},
        {0.0f, 1.0f, 0.111111f, 0.2222f},   {1.0f, -1.0f, 0.0277777f, 0.1111f},
        {-1.0f, -1.0f, 0.0277777f, 0.1111f}, {-1.0f, 1.0f, 0.0277777f, 0.1111f},
        {1.0f, 1.0f, 0.0277777f, 0.1111f},
    };
    static const int32_t inv[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    out->tau = tau;
    out->omega = 1.0f / tau;
    out->fluid_ty = fluid_ty;
    out->soa_offset = soa_offset;
    memcpy(out->e_w_max, ewm, sizeof(ewm));
    for (int i = 0; i < 9; i++)
        for (int k = 0; k < 4; k++) out->inversed_direction[i][k] = inv[i];
}

/* d2q9_node.rs:50 */
float orc_tau_from_viscosity(float viscosity) { return 3.0f * viscosity + 0.5f; }

/* fluid/mod.rs:57-59  glam Vec2::length() = sqrt(x*x + y*y) */
static int is_sd_sphere(float px, float py, float r) {
    float len = sqrtf(px * px + py * py);
    return len <= r;
}

#define OBSTACLE_RADIUS 28.0f /* fluid/mod.rs:1 */

/* fluid/lattice.rs:26-98 */
void orc_init_lattice_material(int32_t nx, int32_t ny, int32_t ty, LatticeInfo *out) {
    float s0x = (float)nx / 7.0f - OBSTACLE_RADIUS, s0y = (float)ny / 2.0f;
    float s1x = (float)nx / 5.0f, s1y = (float)ny / 4.0f;
    float s2x = (float)nx / 5.0f, s2y = (float)ny * 0.75f;
    for (int32_t y = 0; y < ny; y++) {
        for (int32_t x = 0; x < nx; x++) {
            int32_t material = LATTICE_BULK;
            float vx = 0.0f;
            if (ty == FIELD_ANIMATION_CUSTOM) {
                if (x == 0 || x == nx - 1 || y == 0 || y == ny - 1) material = LATTICE_BOUNDARY;
            } else if (ty == FIELD_ANIMATION_LID_DRIVEN_CAVITY) {
                if (x == 0 || x == nx - 1 || y == ny - 1) {
                    material = LATTICE_BOUNDARY;
                } else if (y == 0) {
                    material = LATTICE_GHOST;
                } else if (y == 1) {
                    material = LATTICE_EXTERNAL_FORCE;
                    vx = 0.13f;
                }
            } else if (ty == FIELD_ANIMATION_POISEUILLE) {
                if (y == 0 || y == ny - 1) {
                    material = LATTICE_BOUNDARY;
                } else if (x == 0 || x == nx - 1) {
                    material = LATTICE_GHOST;
                } else if (x == 1) {
                    material = LATTICE_INLET;
                    vx = 0.12f;
                } else if (x == nx - 2) {
                    material = LATTICE_OUTLET;
                } else {
                    float px = (float)x, py = (float)y;
                    if (is_sd_sphere(px - s0x, py - s0y, OBSTACLE_RADIUS) ||
                        is_sd_sphere(px - s1x, py - s1y, OBSTACLE_RADIUS) ||
                        is_sd_sphere(px - s2x, py - s2y, OBSTACLE_RADIUS))
                        material = LATTICE_OBSTACLE;
                }
            }
            LatticeInfo *c = &out[(size_t)y * (size_t)nx + (size_t)x];
            c->material = material;
            c->block_iter = -1;
            c->vx = vx;
            c->vy = 0.0f;
        }
    }
}

static uint64_t splitmix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

/* SURVEY §8d config 5 */
void orc_init_porous_material(int32_t nx, int32_t ny, uint64_t seed, float solid_fraction,
                              LatticeInfo *out) {
    for (int32_t y = 0; y < ny; y++) {
        for (int32_t x = 0; x < nx; x++) {
            int32_t material = LATTICE_BULK;
            float vx = 0.0f;
            if (y == 0 || y == ny - 1) {
                material = LATTICE_BOUNDARY;
            } else if (x == 0 || x == nx - 1) {
                material = LATTICE_GHOST;
            } else if (x == 1) {
                material = LATTICE_INLET;
                vx = 0.12f;
            } else if (x == nx - 2) {
                material = LATTICE_OUTLET;
            } else {
                uint64_t h = splitmix64(seed ^ (((uint64_t)(uint32_t)y << 32) | (uint64_t)(uint32_t)x));
                float r = (float)(h >> 40) / 16777216.0f;
                if (r < solid_fraction) material = LATTICE_OBSTACLE;
            }
            LatticeInfo *c = &out[(size_t)y * (size_t)nx + (size_t)x];
            c->material = material;
            c->block_iter = -1;
            c->vx = vx;
            c->vy = 0.0f;
        }
    }
}

/* ---------------------------------------------------------------- d2q9_fn.wgsl helpers */

static inline int is_boundary(int32_t m) { return m == 2; }   /* d2q9_fn.wgsl:19 */
static inline int is_obstacle(int32_t m) { return m == 4; }   /* d2q9_fn.wgsl:22 */
static inline int is_accelerate(int32_t m) { return m == 3 || m == 6; } /* d2q9_fn.wgsl:24 */

/* ---------------------------------------------------------------- init.wgsl:19-63 */

void orc_init(const LbmUniform *u, int32_t nx, int32_t ny, float *buf0, float *buf1,
              LatticeInfo *info, uint16_t *macro_f16) {
    const size_t N = (size_t)nx * (size_t)ny;
    for (size_t c = 0; c < N; c++) {
        LatticeInfo in = info[c];
        if (is_boundary(in.material) || is_obstacle(in.material)) {
            for (int i = 0; i < 9; i++) {
                buf0[c + (size_t)i * N] = 0.0f;
                buf1[c + (size_t)i * N] = 0.0f;
            }
        } else if (u->fluid_ty == 0) {
            for (int i = 0; i < 9; i++) {
                buf0[c + (size_t)i * N] = u->e_w_max[i][2];
                buf1[c + (size_t)i * N] = 0.0f;
            }
            float temp = u->e_w_max[3][2] * 0.5f;
            buf0[c + 1 * N] = u->e_w_max[1][2] + temp;
            buf0[c + 3 * N] = temp;
            buf1[c + 1 * N] = u->e_w_max[1][2] + temp;
            buf1[c + 3 * N] = temp;
        } else {
            for (int i = 0; i < 9; i++) {
                buf0[c + (size_t)i * N] = u->e_w_max[i][2];
                buf1[c + (size_t)i * N] = 0.0f;
            }
        }
        if (is_accelerate(in.material)) {
            if (in.block_iter > 0) {
                in.block_iter = 0;
                in.material = 1;
                in.vx = 0.0f;
                in.vy = 0.0f;
                info[c] = in;
            }
        }
        if (macro_f16) {
            macro_f16[4 * c + 0] = orc_f32_to_f16(0.0f);
            macro_f16[4 * c + 1] = orc_f32_to_f16(0.0f);
            macro_f16[4 * c + 2] = orc_f32_to_f16(0.0f);
            macro_f16[4 * c + 3] = orc_f32_to_f16(1.0f);
        }
    }
}

/* ---------------------------------------------------------------- collide_stream.wgsl */

/* layout_and_fn.wgsl:38-51 streaming_in — returns the cell index (without the soa plane
 * offset) the pull for `direction` reads. */
static inline size_t streaming_in_cell(const LbmUniform *u, int32_t nx, int32_t ny, int32_t x,
                                       int32_t y, int direction) {
    int inv = u->inversed_direction[direction][0];
    int32_t tx = x + (int32_t)u->e_w_max[inv][0];
    int32_t ty = y + (int32_t)u->e_w_max[inv][1];
    if (tx < 0) tx = nx - 1; else if (tx >= nx) tx = 0;
    if (ty < 0) ty = ny - 1; else if (ty >= ny) ty = 0;
    return (size_t)tx + (size_t)ty * (size_t)nx;
}

/* collide_stream.wgsl:18-22 equilibrium */
static inline float equilibrium(const LbmUniform *u, float vx, float vy, float rho, int i,
                                float usqr) {
    float e_dot_u = u->e_w_max[i][0] * vx + u->e_w_max[i][1] * vy;
    return rho * u->e_w_max[i][2] *
           (1.0f + 3.0f * e_dot_u + 4.5f * (e_dot_u * e_dot_u) - usqr);
}

void orc_collide_stream(const LbmUniform *u, int32_t nx, int32_t ny, const float *rd, float *wr,
                        LatticeInfo *info, uint16_t *macro_f16, float *macro_f32) {
    const size_t N = (size_t)nx * (size_t)ny;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(g_threads)
#endif
    for (int32_t y = 0; y < ny; y++) {
        for (int32_t x = 0; x < nx; x++) {
            const size_t c = (size_t)x + (size_t)y * (size_t)nx; /* d2q9_fn.wgsl:13 */
            LatticeInfo in = info[c];
            if (is_boundary(in.material) || is_obstacle(in.material)) { /* :34-38 */
                if (macro_f16) {
                    for (int k = 0; k < 4; k++) macro_f16[4 * c + k] = 0;
                }
                if (macro_f32) {
                    for (int k = 0; k < 4; k++) macro_f32[4 * c + k] = 0.0f;
                }
                continue;
            }
            float f_i[9];
            float vx = 0.0f, vy = 0.0f, rho = 0.0f;
            for (int i = 0; i < 9; i++) { /* :43-48 */
                f_i[i] = rd[streaming_in_cell(u, nx, ny, x, y, i) + (size_t)i * N];
                rho = rho + f_i[i];
                vx = vx + u->e_w_max[i][0] * f_i[i];
                vy = vy + u->e_w_max[i][1] * f_i[i];
            }
            rho = fminf(fmaxf(rho, 0.8f), 1.2f); /* :49 clamp = min(max(x,lo),hi) */
            vx = vx / rho;                        /* :51 */
            vy = vy / rho;
            float F[9] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (is_accelerate(in.material)) { /* :55-71 */
                if (in.block_iter > 0) {
                    in.block_iter = in.block_iter - 1;
                    if (in.block_iter == 0) in.material = 1;
                }
                info[c] = in;
                float fx = in.vx, fy = in.vy;
                vx = fx * 0.5f / rho;
                vy = fy * 0.5f / rho;
                for (int i = 0; i < 9; i++) {
                    F[i] = u->e_w_max[i][2] * 3.0f * (u->e_w_max[i][0] * fx + u->e_w_max[i][1] * fy);
                }
            }
            if (macro_f16) { /* :74 textureStore -> rgba16float */
                macro_f16[4 * c + 0] = orc_f32_to_f16(vx);
                macro_f16[4 * c + 1] = orc_f32_to_f16(vy);
                macro_f16[4 * c + 2] = orc_f32_to_f16(rho);
                macro_f16[4 * c + 3] = orc_f32_to_f16(1.0f);
            }
            if (macro_f32) {
                macro_f32[4 * c + 0] = vx;
                macro_f32[4 * c + 1] = vy;
                macro_f32[4 * c + 2] = rho;
                macro_f32[4 * c + 3] = 1.0f;
            }
            float usqr = 1.5f * (vx * vx + vy * vy); /* :76 */
            for (int i = 0; i < 9; i++) {             /* :77-87 */
                float temp_val = f_i[i] - u->omega * (f_i[i] - equilibrium(u, vx, vy, rho, i, usqr)) + F[i];
                if (temp_val > u->e_w_max[i][3]) {
                    temp_val = u->e_w_max[i][3];
                } else if (temp_val < 0.0f) {
                    temp_val = 0.0f;
                }
                wr[c + (size_t)i * N] = temp_val;
            }
        }
    }
}

/* ---------------------------------------------------------------- boundary.wgsl:3-35 */

void orc_boundary(const LbmUniform *u, int32_t nx, int32_t ny, float *wr, const LatticeInfo *info) {
    const size_t N = (size_t)nx * (size_t)ny;
    /* The reference runs one invocation per cell concurrently; two adjacent solid cells
     * touch each other's (dead) slots, which is only observable when those slots are
     * non-zero (never, for a mask installed before init).  Serial row-major order when
     * g_threads == 1. */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(g_threads)
#endif
    for (int32_t y = 0; y < ny; y++) {
        for (int32_t x = 0; x < nx; x++) {
            const size_t c = (size_t)x + (size_t)y * (size_t)nx;
            int32_t m = info[c].material;
            if (!(is_boundary(m) || is_obstacle(m))) continue;
            for (int i = 0; i < 9; i++) {
                int32_t qx = x - (int32_t)u->e_w_max[i][0];
                int32_t qy = y - (int32_t)u->e_w_max[i][1];
                if (qx <= 0 || qy <= 0 || qx >= nx - 1 || qy >= ny - 1) continue; /* :19 */
                size_t src = (size_t)qx + (size_t)qy * (size_t)nx + (size_t)i * N;
                float val = wr[src];
                wr[c + (size_t)u->inversed_direction[i][0] * N] = val;
                wr[src] = 0.0f;
            }
        }
    }
}

void orc_step(const LbmUniform *u, int32_t nx, int32_t ny, const float *rd, float *wr,
              LatticeInfo *info, uint16_t *macro_f16, float *macro_f32) {
    orc_collide_stream(u, nx, ny, rd, wr, info, macro_f16, macro_f32);
    orc_boundary(u, nx, ny, wr, info);
}

int orc_step_n(const LbmUniform *u, int32_t nx, int32_t ny, float *buf0, float *buf1,
               LatticeInfo *info, uint16_t *macro_f16, float *macro_f32, int first_swap, int n) {
    int swap = first_swap & 1;
    for (int s = 0; s < n; s++) {
        if (swap == 0)
            orc_step(u, nx, ny, buf0, buf1, info, macro_f16, macro_f32);
        else
            orc_step(u, nx, ny, buf1, buf0, info, macro_f16, macro_f32);
        swap ^= 1;
    }
    return swap;
}

double orc_total_mass(int32_t nx, int32_t ny, const float *buf) {
    const size_t n = (size_t)nx * (size_t)ny * 9u;
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += (double)buf[i];
    return s;
}

/* ---------------------------------------------------------------- particle_update.wgsl */

static inline int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* WGSL f32 -> i32 conversion: truncate toward zero, saturating. */
static inline int32_t f32_to_i32(float v) {
    if (!(v == v)) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int32_t)v;
}

/* particle_update.wgsl:15-21 src_3f */
static inline void src_3f(const FieldUniform *field, const uint16_t *macro_f16, int32_t uu, int32_t vv,
                          float out[3]) {
    int32_t nu = clampi(uu, 0, field->lattice_size[0] - 1);
    int32_t nv = clampi(vv, 0, field->lattice_size[1] - 1);
    size_t c = (size_t)nu + (size_t)nv * (size_t)field->lattice_size[0];
    out[0] = orc_f16_to_f32(macro_f16[4 * c + 0]);
    out[1] = orc_f16_to_f32(macro_f16[4 * c + 1]);
    out[2] = orc_f16_to_f32(macro_f16[4 * c + 2]);
}

/* func/bilinear_interpolate_3f.wgsl:1-12 */
static inline void bilinear_interpolate_3f(const FieldUniform *field, const uint16_t *macro_f16, float uvx,
                                           float uvy, float out[3]) {
    int32_t minX = f32_to_i32(floorf(uvx));
    int32_t minY = f32_to_i32(floorf(uvy));
    float fx = uvx - (float)minX;
    float fy = uvy - (float)minY;
    float a[3], b[3], c[3], d[3];
    src_3f(field, macro_f16, minX, minY, a);
    src_3f(field, macro_f16, minX, minY + 1, b);
    src_3f(field, macro_f16, minX + 1, minY, c);
    src_3f(field, macro_f16, minX + 1, minY + 1, d);
    float wa = (1.0f - fx) * (1.0f - fy);
    float wb = (1.0f - fx) * fy;
    float wc = fx * (1.0f - fy);
    float wd = fx * fy;
    for (int k = 0; k < 3; k++) out[k] = a[k] * wa + b[k] * wb + c[k] * wc + d[k] * wd;
}

void orc_particle_update(const LbmUniform *u, const FieldUniform *field, const ParticleUniform *pu,
                         TrajectoryParticle *particles, Pixel *canvas, const uint16_t *macro_f16) {
    const int poiseuille = (u->fluid_ty == 0);
    for (int32_t gy = 0; gy < pu->num[1]; gy++) {
        for (int32_t gx = 0; gx < pu->num[0]; gx++) {
            size_t p_index = (size_t)gx + (size_t)gy * (size_t)pu->num[0]; /* :28-30 */
            TrajectoryParticle p = particles[p_index];
            if (p.life_time <= 0.1f) { /* :63-66 */
                p.fade = 0.0f;
                p.pos[0] = p.pos_initial[0];
                p.pos[1] = p.pos_initial[1];
                p.life_time = pu->life_time;
            } else {
                p.life_time = p.life_time - 1.0f;
                if (p.fade < 1.0f) {
                    if (p.fade < 0.95f) p.fade = p.fade + 0.1f; else p.fade = 1.0f;
                }
                float ijx = (p.pos[0] / field->lattice_pixel_size[0]) - 0.5f; /* :79 */
                float ijy = (p.pos[1] / field->lattice_pixel_size[1]) - 0.5f;
                float fi[3];
                bilinear_interpolate_3f(field, macro_f16, ijx, ijy, fi);
                p.pos[0] = p.pos[0] + (fi[0] * pu->speed_factor); /* :81 */
                p.pos[1] = p.pos[1] + (fi[1] * pu->speed_factor);
                /* update_canvas :32-53 */
                float speed = fabsf(fi[0]) + fabsf(fi[1]);
                int skip = (!poiseuille && speed < 0.0f) || (poiseuille && speed < 0.015f);
                if (!skip && canvas) {
                    int32_t pcx = f32_to_i32(p.pos[0]);
                    int32_t pcy = f32_to_i32(p.pos[1]);
                    int32_t px = pcx - pu->point_size / 2;
                    int32_t py = pcy - pu->point_size / 2;
                    Pixel pixel = {p.fade, fi[0], fi[1]};
                    for (int32_t dx = 0; dx < pu->point_size; dx++) {
                        for (int32_t dy = 0; dy < pu->point_size; dy++) {
                            int32_t cx = px + dx, cy = py + dy;
                            if (cx >= 0 && cx < field->canvas_size[0] && cy >= 0 && cy < field->canvas_size[1]) {
                                canvas[(size_t)cx + (size_t)field->canvas_size[0] * (size_t)cy] = pixel;
                            }
                        }
                    }
                }
            }
            particles[p_index] = p;
        }
    }
}

/* present.wgsl:19-22,43-49 */
/* curl_update.wgsl:12-33 */
static float orc_tex(const uint16_t *tex, int32_t nx, int32_t ny, int32_t x, int32_t y, int comp) {
    if (x < 0 || x >= nx || y < 0 || y >= ny) return 0.0f; /* out-of-bounds textureLoad: zero (header comment) */
    return orc_f16_to_f32(tex[4 * ((size_t)y * (size_t)nx + (size_t)x) + comp]);
}

void orc_curl_update(int32_t nx, int32_t ny, const uint16_t *macro_f16, uint16_t *curl_f16) {
#pragma omp parallel for schedule(static)
    for (int32_t y = 0; y < ny; y++)
        for (int32_t x = 0; x < nx; x++) {
            const int32_t rx = x + 1 < nx ? x + 1 : nx;   /* min(uv.x + 1, lattice_size.x)  (:21) */
            const int32_t lx = x - 1 > 0 ? x - 1 : 0;     /* max(uv.x - 1, 0)               (:22) */
            const int32_t ty = y - 1 > 0 ? y - 1 : 0;     /* max(uv.y - 1, 0)               (:23) */
            const int32_t by = y + 1 < ny ? y + 1 : ny;   /* min(uv.y + 1, lattice_size.y)  (:24) */
            /* min / max act on both components: y stays (y <= ny), x stays (x <= nx) */
            float curl = orc_tex(macro_f16, nx, ny, rx, y, 1) - orc_tex(macro_f16, nx, ny, lx, y, 1);
            curl = curl + orc_tex(macro_f16, nx, ny, x, ty, 0);
            curl = curl - orc_tex(macro_f16, nx, ny, x, by, 0);
            const size_t c = (size_t)y * (size_t)nx + (size_t)x;
            curl_f16[4 * c + 0] = orc_f32_to_f16(curl * 3.5f + 0.5f); /* -ffp-contract=off: two roundings */
            curl_f16[4 * c + 1] = orc_f32_to_f16(0.0f);
            curl_f16[4 * c + 2] = orc_f32_to_f16(0.0f);
            curl_f16[4 * c + 3] = orc_f32_to_f16(0.0f);
        }
}

/* ---------------------------------------------------------------- colour present of the field
 * assets/wgsl/lbm/present.wgsl:21-46 (+ func/color_space_convert.wgsl:2-9, bufferless.vs.wgsl): the fragment shader of
 * the reference's `render_node` (fluid_simulator.rs:69-87; its draw call is commented out, :243-244).  Per pixel of a
 * canvas_size render target: macro = textureSample(macro_tex, uv), curl = textureSample(curl_tex, uv) through
 * `bilinear_sampler` (util/load_texture.rs:229-243: ClampToEdge, linear), speed = |u.x| + |u.y|,
 * colour = hsv2rgb(curl.x, 0.6 + speed * 1.4, 0.6 + rho * 0.33), alpha = rho.  The unused `canvas[p_index]` /
 * `particle_uniform.color` reads and the dead `angle` (atan2) have no effect on the output.
 * Fragment inputs: position = pixel centre, uv = position.xy / canvas_size.  The filter is the WebGPU formula evaluated
 * in f32 with one rounding per operation (hardware samplers use fixed-point weights of unspecified width, so no
 * implementation is bit-comparable with another; tests/wgsl_ref/runtime.py::textureSample states the same arithmetic). */
static void orc_sample_bilinear(const uint16_t *tex, int32_t nx, int32_t ny, float u, float v, float out[4]) {
    const float cx = u * (float)nx - 0.5f, cy = v * (float)ny - 0.5f;
    const float fx0 = floorf(cx), fy0 = floorf(cy);
    const float fx = cx - fx0, fy = cy - fy0;
    int32_t x0 = (int32_t)fx0, y0 = (int32_t)fy0, x1 = x0 + 1, y1 = y0 + 1;
    x0 = x0 < 0 ? 0 : (x0 > nx - 1 ? nx - 1 : x0);
    x1 = x1 < 0 ? 0 : (x1 > nx - 1 ? nx - 1 : x1);
    y0 = y0 < 0 ? 0 : (y0 > ny - 1 ? ny - 1 : y0);
    y1 = y1 < 0 ? 0 : (y1 > ny - 1 ? ny - 1 : y1);
    const float gx = 1.0f - fx, gy = 1.0f - fy;
    const float w00 = gx * gy, w10 = fx * gy, w01 = gx * fy, w11 = fx * fy;
    for (int k = 0; k < 4; k++) {
        float s = w00 * orc_tex(tex, nx, ny, x0, y0, k);
        s = s + w10 * orc_tex(tex, nx, ny, x1, y0, k);
        s = s + w01 * orc_tex(tex, nx, ny, x0, y1, k);
        s = s + w11 * orc_tex(tex, nx, ny, x1, y1, k);
        out[k] = s;
    }
}

/* func/color_space_convert.wgsl:2-9.  K = (1, 2/3, 1/3, 3); fract(e) = e - floor(e); mix(a, b, t) = a*(1-t) + b*t */
static void orc_hsv2rgb(float h, float s, float v, float rgb[3]) {
    const float K[3] = {1.0f, 2.0f / 3.0f, 1.0f / 3.0f};
    for (int k = 0; k < 3; k++) {
        const float t = h + K[k];
        const float p = fabsf((t - floorf(t)) * 6.0f - 3.0f);
        float c = p - 1.0f;
        c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c); /* clamp = min(max(c, 0), 1) */
        rgb[k] = v * (1.0f * (1.0f - s) + c * s);
    }
}

void orc_present(const FieldUniform *field, const uint16_t *macro_f16, const uint16_t *curl_f16, int32_t row0,
                 int32_t rows, float *out_rgba) {
    const int32_t nx = field->lattice_size[0], ny = field->lattice_size[1];
    const int32_t W = field->canvas_size[0], H = field->canvas_size[1];
#pragma omp parallel for schedule(static)
    for (int32_t py = row0; py < row0 + rows; py++)
        for (int32_t px = 0; px < W; px++) {
            const float u = ((float)px + 0.5f) / (float)W, v = ((float)py + 0.5f) / (float)H;
            float m[4], c[4], rgb[3];
            orc_sample_bilinear(macro_f16, nx, ny, u, v, m);
            orc_sample_bilinear(curl_f16, nx, ny, u, v, c);
            const float speed = fabsf(m[0]) + fabsf(m[1]);
            orc_hsv2rgb(c[0], 0.6f + speed * 1.4f, 0.6f + m[2] * 0.33f, rgb);
            float *o = out_rgba + 4 * ((size_t)(py - row0) * (size_t)W + (size_t)px);
            o[0] = rgb[0]; o[1] = rgb[1]; o[2] = rgb[2]; o[3] = m[2];
        }
}

void orc_canvas_fade(const FieldUniform *field, const ParticleUniform *pu, Pixel *canvas) {
    const size_t n = (size_t)field->canvas_size[0] * (size_t)field->canvas_size[1];
    for (size_t i = 0; i < n; i++) {
        Pixel p = canvas[i];
        if (p.alpha > 0.001f) {
            if (p.alpha >= 0.2f) p.alpha = p.alpha * pu->fade_out_factor;
            else p.alpha = p.alpha * 0.5f;
            canvas[i] = p;
        }
    }
}

/* ---------------------------------------------------------------- host mutations */

/* d2q9_node.rs:215-245 */
size_t orc_add_obstacle(int32_t nx, int32_t ny, LatticeInfo *mirror, uint32_t x, uint32_t y,
                        LatticeInfo *patch, uint64_t *byte_offset) {
    (void)ny;
    LatticeInfo obstacle = {LATTICE_OBSTACLE, -1, 0.0f, 0.0f};
    float cx = (float)x + 0.5f, cy = (float)y + 0.5f;
    uint32_t min_y = y - (uint32_t)OBSTACLE_RADIUS;
    uint32_t max_y = min_y + (uint32_t)OBSTACLE_RADIUS * 2u;
    size_t n = 0;
    for (uint32_t yy = min_y; yy < max_y; yy++) {
        for (uint32_t xx = 0; xx < (uint32_t)nx; xx++) {
            size_t index = (size_t)nx * yy + xx;
            if (is_sd_sphere(((float)xx + 0.5f) - cx, ((float)yy + 0.5f) - cy, OBSTACLE_RADIUS)) {
                mirror[index] = obstacle;
                patch[n++] = obstacle;
            } else {
                patch[n++] = mirror[index];
            }
        }
    }
    *byte_offset = (uint64_t)((uint32_t)nx * min_y) * 16u;
    return n;
}

/* Rust `f32 as u32`: saturating, NaN -> 0 */
static inline uint32_t f32_as_u32(float v) {
    if (!(v == v) || v <= 0.0f) return 0u;
    if (v >= 4294967296.0f) return 4294967295u;
    return (uint32_t)v;
}

/* fluid_simulator.rs:137-152 */
int orc_on_click_guard(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, float px, float py,
                       uint32_t *x, uint32_t *y) {
    if (px <= 0.0f || py <= 0.0f) return 0;
    uint32_t cx = f32_as_u32(px) / lattice_pixel_size;
    uint32_t cy = f32_as_u32(py) / lattice_pixel_size;
    uint32_t half = (uint32_t)OBSTACLE_RADIUS;
    if (cx < half || cx >= (uint32_t)nx - (half + 2) || cy < half || cy >= (uint32_t)ny - (half + 2)) return 0;
    *x = cx;
    *y = cy;
    return 1;
}

/* d2q9_node.rs:263-300 */
size_t orc_add_external_force(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, float pos_x,
                              float pos_y, float pre_x, float pre_y, uint64_t *byte_offsets,
                              LatticeInfo *cells, size_t cap) {
    float ddx = pos_x - pre_x, ddy = pos_y - pre_y;
    float dis = sqrtf(ddx * ddx + ddy * ddy); /* glam distance */
    float force = 0.1f * (dis / 20.0f);
    if (force > 0.12f) force = 0.12f;
    float ridian = atan2f(pos_y - pre_y, pos_x - pre_x);
    float vx = force * cosf(ridian);
    float vy = force * sinf(ridian);
    LatticeInfo cell = {LATTICE_EXTERNAL_FORCE, 90, vx, vy};
    float c = ceilf(dis / (float)(lattice_pixel_size - 1));
    float step = dis / c;
    size_t n = 0;
    int32_t ci = f32_to_i32(c);
    for (int32_t i = 0; i < ci; i++) {
        float d = step * (float)i;
        float qx = roundf(pre_x + d * cosf(ridian));
        float qy = roundf(pre_y + d * sinf(ridian));
        uint32_t x = f32_as_u32(qx) / lattice_pixel_size;
        uint32_t y = f32_as_u32(qy) / lattice_pixel_size;
        if (x < 1 || x >= (uint32_t)nx - 2 || y < 1 || y >= (uint32_t)ny - 2) continue;
        if (n < cap) {
            byte_offsets[n] = (uint64_t)((uint32_t)nx * y + x) * 16u;
            cells[n] = cell;
        }
        n++;
    }
    return n;
}

/* d2q9_node.rs:61-76 */
void orc_field_uniform_new(int32_t nx, int32_t ny, uint32_t lattice_pixel_size, int32_t canvas_w,
                           int32_t canvas_h, FieldUniform *out) {
    memset(out, 0, sizeof(*out));
    out->lattice_size[0] = nx;
    out->lattice_size[1] = ny;
    out->lattice_pixel_size[0] = (float)lattice_pixel_size;
    out->lattice_pixel_size[1] = (float)lattice_pixel_size;
    out->canvas_size[0] = canvas_w;
    out->canvas_size[1] = canvas_h;
    /* (_, sx, sy) = fullscreen_factor((canvas_w, canvas_h), fovy), util/matrix_helper.rs:19-41: the longer side over
     * the shorter one on its axis, 1 on the other; proj_ratio = [sx, sy], ndc_pixel = [sx * 2.0 / w, sy * 2.0 / h] */
    {
        float vw = (float)canvas_w, vh = (float)canvas_h, sx = 1.0f, sy = 1.0f;
        if (vh > vw) sy = vh / vw; else sx = vw / vh;
        out->proj_ratio[0] = sx;
        out->proj_ratio[1] = sy;
        out->ndc_pixel[0] = sx * 2.0f / vw;
        out->ndc_pixel[1] = sy * 2.0f / vh;
    }
    out->speed_ty = 1;
}

/* lib.rs:247-264 */
void orc_particle_grid(uint32_t canvas_w, uint32_t canvas_h, int32_t count, int32_t *num_x,
                       int32_t *num_y) {
    float ratio = (float)canvas_w / (float)canvas_h;
    float x = ceilf(sqrtf((float)count * ratio));
    *num_x = (int32_t)f32_as_u32(x);
    *num_y = (int32_t)f32_as_u32(ceilf(x * (1.0f / ratio)));
}

static inline float unit_from(uint64_t *state) {
    *state = *state + 0x9e3779b97f4a7c15ull;
    uint64_t z = *state;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) / 16777216.0f; /* [0,1) */
}

/* lib.rs:275-316 with a seeded stream */
void orc_init_trajectory_particles(uint32_t canvas_w, uint32_t canvas_h, int32_t num_x,
                                   int32_t num_y, float life_time, uint64_t seed,
                                   TrajectoryParticle *out) {
    uint64_t st = seed;
    float step_x = (float)canvas_w / (float)(num_x - 1);
    float step_y = (float)canvas_h / (float)(num_y - 1);
    float life_hi = (life_time <= 0.0f) ? 1.0f : life_time;
    size_t n = 0;
    for (int32_t x = 0; x < num_x; x++) {
        float pixel_x = step_x * (float)x;
        for (int32_t y = 0; y < num_y; y++) {
            float jx = -step_x + (step_x - (-step_x)) * unit_from(&st);
            float jy = -step_y + (step_y - (-step_y)) * unit_from(&st);
            TrajectoryParticle p;
            p.pos[0] = pixel_x + jx;
            p.pos[1] = step_y * (float)y + jy;
            if (life_time <= 1.0f) {
                p.pos_initial[0] = 0.0f + (step_x - 0.0f) * unit_from(&st);
                p.pos_initial[1] = p.pos[1];
                p.life_time = 0.0f;
            } else {
                p.pos_initial[0] = p.pos[0];
                p.pos_initial[1] = p.pos[1];
                p.life_time = 0.0f + (life_hi - 0.0f) * unit_from(&st);
            }
            p.fade = 0.0f;
            out[n++] = p;
        }
    }
}
