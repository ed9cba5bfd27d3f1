/*
 * lbm_wire.h — byte layouts that cross the drop-in boundary of the D2Q9 LBM path.
 *
 * Every struct here is the little-endian #[repr(C)] / bytemuck::Pod image of the
 * reference's Rust struct of the same name (paths relative to the reference tree):
 *
 *   LbmUniform          simuverse/src/fluid/mod.rs:12-29        (304 B)
 *   FieldUniform        simuverse/src/lib.rs:161-179            ( 48 B)
 *   LatticeInfo         simuverse/src/fluid/lattice.rs:5-13     ( 16 B)
 *   ParticleUniform     simuverse/src/lib.rs:180-196            ( 48 B)
 *   TrajectoryParticle  simuverse/src/lib.rs:198-205            ( 24 B)
 *   Pixel               simuverse/src/lib.rs:235-243            ( 12 B)
 *
 * The WGSL views of the same bytes are assets/wgsl/lbm/struct/lbm_uniform.wgsl:2-12,
 * lbm/struct/lattice_info.wgsl:2-8, struct/field.wgsl:2-16, struct/particle.wgsl,
 * struct/pixel.wgsl.
 */
#ifndef LBM_WIRE_H
#define LBM_WIRE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* D2Q9 direction numbering (fluid/mod.rs:22-25), y grows downward:
 *   6 2 5
 *   3 0 1
 *   7 4 8
 */
#define LBM_Q 9

typedef struct LbmUniform {
    float   tau;
    float   omega;                      /* 1.0f / tau, computed on the host in f32 (fluid/mod.rs:35) */
    int32_t fluid_ty;                   /* 0: Poiseuille / Custom, 1: LidDrivenCavity (d2q9_node.rs:53-57) */
    int32_t soa_offset;                 /* nx*ny in the reference; informational here (64-bit indexing inside) */
    float   e_w_max[LBM_Q][4];          /* (e.x, e.y, weight, per-direction max) */
    int32_t inversed_direction[LBM_Q][4]; /* value replicated x4, shader reads .x */
} LbmUniform;

typedef struct FieldUniform {
    int32_t lattice_size[2];
    float   lattice_pixel_size[2];
    int32_t canvas_size[2];
    float   proj_ratio[2];
    float   ndc_pixel[2];
    int32_t speed_ty;                   /* 1 for the LBM simulator (d2q9_node.rs:74) */
    float   _padding;
} FieldUniform;

typedef struct LatticeInfo {
    int32_t material;                   /* LatticeType */
    int32_t block_iter;                 /* countdown for transient force cells; -1 = static */
    float   vx;
    float   vy;
} LatticeInfo;

/* fluid/lattice.rs:15-24 */
enum LatticeType {
    LATTICE_BULK = 1,
    LATTICE_BOUNDARY = 2,
    LATTICE_INLET = 3,
    LATTICE_OBSTACLE = 4,
    LATTICE_OUTLET = 5,
    LATTICE_EXTERNAL_FORCE = 6,
    LATTICE_GHOST = 7
};

/* lib.rs:119-141 — only the three LBM presets matter to this path */
enum FieldAnimationType {
    FIELD_ANIMATION_POISEUILLE = 4,
    FIELD_ANIMATION_LID_DRIVEN_CAVITY = 5,
    FIELD_ANIMATION_CUSTOM = 6
};

typedef struct ParticleUniform {
    float   color[4];
    int32_t num[2];
    int32_t point_size;
    float   life_time;
    float   fade_out_factor;
    float   speed_factor;
    int32_t color_ty;
    int32_t is_only_update_pos;
} ParticleUniform;

typedef struct TrajectoryParticle {
    float pos[2];
    float pos_initial[2];
    float life_time;
    float fade;
} TrajectoryParticle;

typedef struct Pixel {
    float alpha;
    float velocity_x;
    float velocity_y;
} Pixel;

#ifdef __cplusplus
}
static_assert(sizeof(LbmUniform) == 304, "LbmUniform wire size");
static_assert(sizeof(FieldUniform) == 48, "FieldUniform wire size");
static_assert(sizeof(LatticeInfo) == 16, "LatticeInfo wire size");
static_assert(sizeof(ParticleUniform) == 48, "ParticleUniform wire size");
static_assert(sizeof(TrajectoryParticle) == 24, "TrajectoryParticle wire size");
static_assert(sizeof(Pixel) == 12, "Pixel wire size");
#else
_Static_assert(sizeof(LbmUniform) == 304, "LbmUniform wire size");
_Static_assert(sizeof(FieldUniform) == 48, "FieldUniform wire size");
_Static_assert(sizeof(LatticeInfo) == 16, "LatticeInfo wire size");
_Static_assert(sizeof(ParticleUniform) == 48, "ParticleUniform wire size");
_Static_assert(sizeof(TrajectoryParticle) == 24, "TrajectoryParticle wire size");
_Static_assert(sizeof(Pixel) == 12, "Pixel wire size");
#endif

#endif /* LBM_WIRE_H */
