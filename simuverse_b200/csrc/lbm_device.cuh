// lbm_device.cuh — device-side data model and per-cell arithmetic of the D2Q9 step.
//
// Arithmetic contract (collide_stream.wgsl:43-87): every f32 operation
// below is an explicit round-to-nearest intrinsic (__fadd_rn/__fmul_rn/__fdiv_rn), which nvcc
// never contracts into FMA, in the reference's source order.  Sub-expressions whose e_i
// component is 0 are dropped and opposite directions share eu, eu*eu: both are exact
// identities for the finite, non-negative distributions the step produces (adding a
// zero-signed product changes nothing, -(a*b) == (-a)*b under RN).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lbm_wire.h"

namespace lbm {

// cell class plane (1 B/cell): what the hot kernel needs to know about a cell
enum : uint8_t {
    CLS_FLUID = 0,     // plain fluid, no solid neighbour to bounce into -> fast path
    CLS_FLUID_NB = 1,  // fluid with >=1 solid neighbour (consult the nbr plane)
    CLS_SOLID = 2,     // material 2 (Boundary) or 4 (Obstacle): skipped (collide_stream.wgsl:34)
    CLS_ACCEL = 3,     // material 3 (Inlet) or 6 (ExternalForce): d2q9_fn.wgsl:24
    CLS_FLIPPED = 4,   // force cell whose countdown hit 0 in the step just executed: it was still
                       // forced in that step (what the on-demand field read must reproduce) and is
                       // plain fluid from the next step on, which also retires this state
};

// D2Q9 lattice vectors (fluid/mod.rs:39-52), fixed; lbm_write_uniform rejects anything else.
// constexpr functions rather than __constant__ tables: under full unrolling they fold to immediates
__host__ __device__ constexpr int dir_ex(int i) { return (i == 1 || i == 5 || i == 8) ? 1 : ((i == 3 || i == 6 || i == 7) ? -1 : 0); }
__host__ __device__ constexpr int dir_ey(int i) { return (i == 4 || i == 7 || i == 8) ? 1 : ((i == 2 || i == 5 || i == 6) ? -1 : 0); }
__host__ __device__ constexpr int dir_inv(int i) { return i == 0 ? 0 : (i <= 2 ? i + 2 : (i <= 4 ? i - 2 : (i <= 6 ? i + 2 : i - 2))); }
static_assert(dir_inv(1) == 3 && dir_inv(2) == 4 && dir_inv(3) == 1 && dir_inv(4) == 2 && dir_inv(5) == 7 &&
                  dir_inv(6) == 8 && dir_inv(7) == 5 && dir_inv(8) == 6, "fluid/mod.rs:50-52");
static_assert(dir_ex(5) == 1 && dir_ey(5) == -1 && dir_ex(6) == -1 && dir_ey(6) == -1 && dir_ex(7) == -1 &&
                  dir_ey(7) == 1 && dir_ex(8) == 1 && dir_ey(8) == 1 && dir_ex(2) == 0 && dir_ey(2) == -1 && dir_ey(4) == 1,
              "fluid/mod.rs:39-49");

struct Coef {
    float omega;
    float w[9];
    float mx[9];
    int fluid_ty;
};

// Everything a kernel needs to know about one y-slab.  Passed by value (__grid_constant__).
struct SlabParams {
    int nx;            // lattice width
    int ny;            // GLOBAL lattice height
    int y0;            // global row of local row 0
    int h;             // rows owned by this slab
    int pitch;         // row pitch of f / cls / nbr planes, in elements (multiple of 32)
    size_t plane;      // plane stride of the own distribution buffers, in floats
    float *f[2];       // own distributions: f[b][dir*plane + l*pitch + x]
    // Row y0-1 ("up") and row y0+h ("dn") of each buffer: own memory when world==1 (periodic
    // wrap, layout_and_fn.wgsl:45-49), otherwise the neighbour slab's memory mapped over NVLink.
    float *up[2];
    float *dn[2];
    size_t up_plane;   // plane stride of the memory `up` points into
    size_t dn_plane;
    uint8_t *cls;      // [l*pitch + x]
    const uint8_t *cls_up; // class row of row y0-1 / y0+h (own wrapped rows when world==1, else a local copy
    const uint8_t *cls_dn; // derived from the info halo rows); read by the two-update kernel only
    uint8_t *nbr;      // fluid cell: bit (i-1) set = strictly interior and cell+e_i is solid (bounce);
                       // solid cell: bit (k-1) set = slot k is dead (its only reader is solid too)
    LatticeInfo *info; // (h+2) rows of nx: row 0 = halo y0-1, rows 1..h owned, row h+1 = halo
    __half *macro16;   // h*nx texels of 4 halfs, or nullptr
    __half *macro16_mid; // two-update sweeps with tracer particles: the texture update 1 stores into (lbm_fused.cuh)
    float *macro32;    // 3 planes of h*nx f32 (u.x,u.y,rho), or nullptr
    Coef k;
};

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// collide_stream.wgsl:79-83: `if t > max {max} else if t < 0 {0}` (a NaN passes through)
__device__ __forceinline__ float clamp_dir(float t, float mx) {
    return t > mx ? mx : (t < 0.0f ? 0.0f : t);
}

// collide_stream.wgsl:43-51 moments of the pulled distributions.
__device__ __forceinline__ void moments(const float (&f)[9], float &rho, float &ux, float &uy) {
    float r = fadd(0.0f, f[0]);
    r = fadd(r, f[1]); r = fadd(r, f[2]); r = fadd(r, f[3]); r = fadd(r, f[4]);
    r = fadd(r, f[5]); r = fadd(r, f[6]); r = fadd(r, f[7]); r = fadd(r, f[8]);
    // u.x = sum e_i.x f_i in i order: +f1 -f3 +f5 -f6 -f7 +f8 ; u.y: -f2 +f4 -f5 -f6 +f7 +f8
    float x = fadd(0.0f, f[1]);
    x = fsub(x, f[3]); x = fadd(x, f[5]); x = fsub(x, f[6]); x = fsub(x, f[7]); x = fadd(x, f[8]);
    float y = fsub(0.0f, f[2]);
    y = fadd(y, f[4]); y = fsub(y, f[5]); y = fsub(y, f[6]); y = fadd(y, f[7]); y = fadd(y, f[8]);
    r = fminf(fmaxf(r, 0.8f), 1.2f); // :49
    rho = r;
    ux = fdiv(x, r);                 // :51
    uy = fdiv(y, r);
}

// BGK relaxation toward equilibrium for a pair of opposite directions p (eu = a) and m (eu = -a).
__device__ __forceinline__ void relax_pair(const Coef &k, float rho, float usqr, float a, int p, int m,
                                           float (&f)[9]) {
    const float c3 = fmul(3.0f, a);
    const float c45 = fmul(4.5f, fmul(a, a));
    const float feq_p = fmul(fmul(rho, k.w[p]), fsub(fadd(fadd(1.0f, c3), c45), usqr));
    const float feq_m = fmul(fmul(rho, k.w[m]), fsub(fadd(fsub(1.0f, c3), c45), usqr));
    f[p] = fsub(f[p], fmul(k.omega, fsub(f[p], feq_p)));
    f[m] = fsub(f[m], fmul(k.omega, fsub(f[m], feq_m)));
}

// collide_stream.wgsl:76-87 for a non-accelerate cell (F_i = 0; the `+ 0.0` is dropped: it
// only turns -0 into +0, and t is never -0 because f_i never is).  In place: f -> post-collision.
__device__ __forceinline__ void collide_plain(const Coef &k, float rho, float ux, float uy, float (&f)[9]) {
    const float usqr = fmul(1.5f, fadd(fmul(ux, ux), fmul(uy, uy)));
    {   // direction 0: eu = 0
        const float feq = fmul(fmul(rho, k.w[0]), fsub(1.0f, usqr));
        f[0] = fsub(f[0], fmul(k.omega, fsub(f[0], feq)));
    }
    relax_pair(k, rho, usqr, ux, 1, 3, f);
    relax_pair(k, rho, usqr, uy, 4, 2, f);
    relax_pair(k, rho, usqr, fsub(ux, uy), 5, 7, f);
    relax_pair(k, rho, usqr, fadd(ux, uy), 8, 6, f);
#pragma unroll
    for (int i = 0; i < 9; i++) f[i] = clamp_dir(f[i], k.mx[i]);
}

// collide_stream.wgsl:64-87 for an accelerate cell, written out literally (rare path).
__device__ __forceinline__ void collide_forced(const Coef &k, float rho, float ux, float uy, float fx, float fy,
                                               float (&f)[9]) {
    const float usqr = fmul(1.5f, fadd(fmul(ux, ux), fmul(uy, uy)));
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const float ex = (float)dir_ex(i), ey = (float)dir_ey(i);
        const float eu = fadd(fmul(ex, ux), fmul(ey, uy));
        const float feq = fmul(fmul(rho, k.w[i]),
                               fsub(fadd(fadd(1.0f, fmul(3.0f, eu)), fmul(4.5f, fmul(eu, eu))), usqr));
        const float Fi = fmul(fmul(k.w[i], 3.0f), fadd(fmul(ex, fx), fmul(ey, fy)));
        const float t = fadd(fsub(f[i], fmul(k.omega, fsub(f[i], feq))), Fi);
        f[i] = clamp_dir(t, k.mx[i]);
    }
}

struct RowRef {
    float *p;      // plane 0 of the row
    size_t plane;  // plane stride
};

// Row l of buffer b; l == -1 and l == h resolve to the neighbour rows.
__device__ __forceinline__ RowRef row_ref(const SlabParams &P, int b, int l) {
    RowRef r;
    if (l < 0) { r.p = P.up[b]; r.plane = P.up_plane; }
    else if (l >= P.h) { r.p = P.dn[b]; r.plane = P.dn_plane; }
    else { r.p = P.f[b] + (size_t)l * P.pitch; r.plane = P.plane; }
    return r;
}

__device__ __forceinline__ void store_macro(const SlabParams &P, int x, int l, float ux, float uy, float rho,
                                            float one) {
    const size_t c = (size_t)l * P.nx + x;
    if (P.macro16) {
        __half2 a = __halves2half2(__float2half_rn(ux), __float2half_rn(uy));
        __half2 b = __halves2half2(__float2half_rn(rho), __float2half_rn(one));
        uint2 v;
        v.x = *reinterpret_cast<uint32_t *>(&a);
        v.y = *reinterpret_cast<uint32_t *>(&b);
        reinterpret_cast<uint2 *>(P.macro16)[c] = v;
    }
    if (P.macro32) {
        const size_t n = (size_t)P.h * P.nx;
        P.macro32[c] = ux;
        P.macro32[n + c] = uy;
        P.macro32[2 * n + c] = rho;
    }
}

// Accelerate-cell bookkeeping (collide_stream.wgsl:55-66): countdown, material flip, info
// write-back; returns the force and keeps the class plane coherent with the flip.
__device__ __forceinline__ void accel_update(const SlabParams &P, int x, int l, float &fx, float &fy) {
    LatticeInfo *ip = P.info + (size_t)(l + 1) * P.nx + x;
    LatticeInfo in = *ip;
    if (in.block_iter > 0) {
        in.block_iter -= 1;
        if (in.block_iter == 0) {
            in.material = 1;
            P.cls[(size_t)l * P.pitch + x] = CLS_FLIPPED;
        }
    }
    *ip = in;
    fx = in.vx;
    fy = in.vy;
}

// Solid cell: write the zeros the reference holds in slots nobody reads (see k_derive), so the
// destination buffer is written in whole sectors.  Slot 0 of a strictly interior solid is zeroed by
// boundary.wgsl:28-31 (i = 0: n = cell itself) as well.
__device__ __forceinline__ void zero_dead_slots(const SlabParams &P, float *wc, uint32_t dead, int x, int y) {
    if (x > 0 && x < P.nx - 1 && y > 0 && y < P.ny - 1) wc[0] = 0.0f;
#pragma unroll
    for (int k = 1; k < 9; k++)
        if ((dead >> (k - 1)) & 1u) wc[(size_t)k * P.plane] = 0.0f;
}

// Moments, inlet / force handling, BGK collision and the macro store of one non-solid cell whose nine
// distributions have been pulled into f (collide_stream.wgsl:43-87 after the pull).  In place.
template <int MODE>
__device__ __forceinline__ void collide_cell(const SlabParams &P, uint8_t c, uint8_t nb, int x, int l, float (&f)[9]) {
    const size_t cl = (size_t)l * P.pitch + x;
    float rho, ux, uy;
    moments(f, rho, ux, uy);
    if (MODE == 0 && c == CLS_FLIPPED) P.cls[cl] = nb ? CLS_FLUID_NB : CLS_FLUID;
    if (c == CLS_ACCEL || (MODE == 1 && c == CLS_FLIPPED)) {
        float fx, fy;
        if (MODE == 0) {
            accel_update(P, x, l, fx, fy);
        } else {
            const LatticeInfo in = P.info[(size_t)(l + 1) * P.nx + x];
            fx = in.vx;
            fy = in.vy;
        }
        ux = fdiv(fmul(fx, 0.5f), rho);             // :66
        uy = fdiv(fmul(fy, 0.5f), rho);
        if (MODE == 0) collide_forced(P.k, rho, ux, uy, fx, fy, f);
    } else if (MODE == 0) {
        collide_plain(P.k, rho, ux, uy, f);
    }
    if (MODE == 1 || P.macro16 || P.macro32) store_macro(P, x, l, ux, uy, rho, 1.0f);
}

// One cell of the fused step, generic in every respect: periodic wrap in x, neighbour rows
// in y, bounce-back scatter, accelerate cells.  MODE: 0 = full step, 1 = macro only (no
// collision/stores/info mutation; used by the on-demand field read).
//
// Bounce-back (boundary.wgsl:16-33) is fused as a scatter: if this cell is strictly interior
// and cell+e_i is solid, the post-collision f_i goes to slot inv(i) of that solid cell and 0
// to the own slot — exactly the state the reference's second pass leaves.
template <int MODE>
__device__ __forceinline__ void update_cell(const SlabParams &P, int rb, int x, int l) {
    const size_t cl = (size_t)l * P.pitch + x;
    const uint8_t c = P.cls[cl];
    if (c == CLS_SOLID) {
        if (MODE == 1 || P.macro16 || P.macro32) store_macro(P, x, l, 0.0f, 0.0f, 0.0f, 0.0f);
        if (MODE == 0) zero_dead_slots(P, P.f[rb ^ 1] + cl, P.nbr[cl], x, P.y0 + l);
        return;
    }
    const int xm = (x == 0) ? P.nx - 1 : x - 1;      // layout_and_fn.wgsl:40-44
    const int xp = (x == P.nx - 1) ? 0 : x + 1;
    const RowRef r0 = row_ref(P, rb, l);
    const RowRef ru = row_ref(P, rb, l - 1);        // source row for e_y = +1
    const RowRef rd = row_ref(P, rb, l + 1);        // source row for e_y = -1
    float f[9];
    f[0] = r0.p[x];
    f[1] = r0.p[1 * r0.plane + xm];
    f[2] = rd.p[2 * rd.plane + x];
    f[3] = r0.p[3 * r0.plane + xp];
    f[4] = ru.p[4 * ru.plane + x];
    f[5] = rd.p[5 * rd.plane + xm];
    f[6] = rd.p[6 * rd.plane + xp];
    f[7] = ru.p[7 * ru.plane + xp];
    f[8] = ru.p[8 * ru.plane + xm];
    const uint8_t nb = (c == CLS_FLUID) ? 0 : P.nbr[cl];
    collide_cell<MODE>(P, c, nb, x, l, f);
    if (MODE == 1) return;

    const int wb = rb ^ 1;
    float *w0 = P.f[wb] + cl;
    if (nb == 0) {
#pragma unroll
        for (int i = 0; i < 9; i++) w0[(size_t)i * P.plane] = f[i];
        return;
    }
    w0[0] = f[0];
#pragma unroll
    for (int i = 1; i < 9; i++) {
        if ((nb >> (i - 1)) & 1) {
            const RowRef rt = row_ref(P, wb, l + dir_ey(i));
            rt.p[(size_t)dir_inv(i) * rt.plane + (x + dir_ex(i))] = f[i];
            w0[(size_t)i * P.plane] = 0.0f;
        } else {
            w0[(size_t)i * P.plane] = f[i];
        }
    }
}

}  // namespace lbm
