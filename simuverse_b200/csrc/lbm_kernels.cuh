// lbm_kernels.cuh — kernels of the D2Q9 path other than the vectorised step
// (which lives in lbm_step_vec.cuh): generic step, init, class/neighbour derivation,
// device-side preset generation, mass reduction, on-demand macro field.
#pragma once

#include "lbm_device.cuh"

namespace lbm {

// ------------------------------------------------------------------ generic step
// One thread per cell over rows [l0, l1). Serves as the A/B-testing baseline
// (LBM_FLAG_KERNEL_GENERIC) and as the on-demand macro pass (MODE 1).
template <int MODE>
__global__ void __launch_bounds__(256) k_step_generic(const __grid_constant__ SlabParams P, int rb, int l0, int l1) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = l0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.nx || l >= l1) return;
    update_cell<MODE>(P, rb, x, l);
}

// ------------------------------------------------------------------ init.wgsl:19-63
__global__ void __launch_bounds__(256) k_init(const __grid_constant__ SlabParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.nx || l >= P.h) return;
    LatticeInfo *ip = P.info + (size_t)(l + 1) * P.nx + x;
    LatticeInfo in = *ip;
    const size_t cl = (size_t)l * P.pitch + x;
    float *b0 = P.f[0] + cl, *b1 = P.f[1] + cl;
    const bool two = P.f[1] != P.f[0]; // the in-place (AA) variant keeps only collide_cell = buffer 0
    const bool solid = in.material == 2 || in.material == 4;
    if (solid) {
#pragma unroll
        for (int i = 0; i < 9; i++) { b0[i * P.plane] = 0.0f; if (two) b1[i * P.plane] = 0.0f; }
    } else {
#pragma unroll
        for (int i = 0; i < 9; i++) { b0[i * P.plane] = P.k.w[i]; if (two) b1[i * P.plane] = 0.0f; }
        if (P.k.fluid_ty == 0) { // isPoiseuilleFlow(): bias along +x (init.wgsl:39-43)
            const float temp = fmul(P.k.w[3], 0.5f);
            const float f1 = fadd(P.k.w[1], temp);
            b0[1 * P.plane] = f1; b0[3 * P.plane] = temp;
            if (two) { b1[1 * P.plane] = f1; b1[3 * P.plane] = temp; }
        }
    }
    if ((in.material == 3 || in.material == 6) && in.block_iter > 0) { // init.wgsl:51-59
        in.block_iter = 0; in.material = 1; in.vx = 0.0f; in.vy = 0.0f;
        *ip = in;
    }
    if (P.macro16 || P.macro32) store_macro(P, x, l, 0.0f, 0.0f, 0.0f, 1.0f); // init.wgsl:62
}

// ------------------------------------------------------------------ class / neighbour planes
// Derived from the authoritative LatticeInfo buffer for owned rows [l0, l1).
// armed (optional): receives the largest block_iter of the inlet / force cells seen (atomicMax) — the host
// keeps the two-update kernel off while a countdown is running (collide_stream.wgsl:55-62).
__global__ void __launch_bounds__(256) k_derive(const __grid_constant__ SlabParams P, int l0, int l1, unsigned int *armed) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = l0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.nx || l >= l1) return;
    const LatticeInfo *row = P.info + (size_t)(l + 1) * P.nx;
    const int m = row[x].material;
    if (armed && (m == 3 || m == 6) && row[x].block_iter > 0) atomicMax(armed, (unsigned int)row[x].block_iter);
    const int y = P.y0 + l;
    uint8_t nb = 0;
    const bool solid = (m == 2 || m == 4);
    if (solid) {
        // For a solid cell the byte lists its DEAD slots: slot k is only ever read by the cell at
        // wrap(cell + e_k); if that cell is solid too, nobody reads the slot and the reference holds 0
        // there (boundary.wgsl:28-31 moves zeros between adjacent solids).  The step writes those zeros
        // so that whole 32-byte sectors of the destination buffer get written (no DRAM fill reads).
#pragma unroll
        for (int i = 1; i < 9; i++) {
            int xx = x + dir_ex(i);
            if (xx < 0) xx = P.nx - 1; else if (xx >= P.nx) xx = 0;
            const int mm = row[(ptrdiff_t)dir_ey(i) * P.nx + xx].material; // halo rows hold the wrapped rows
            if (mm == 2 || mm == 4) nb |= (uint8_t)(1u << (i - 1));
        }
    } else if (x > 0 && x < P.nx - 1 && y > 0 && y < P.ny - 1) {
        // boundary.wgsl:19 — only strictly interior cells ever receive a bounce-back
#pragma unroll
        for (int i = 1; i < 9; i++) {
            const int mm = row[(ptrdiff_t)dir_ey(i) * P.nx + x + dir_ex(i)].material;
            if (mm == 2 || mm == 4) nb |= (uint8_t)(1u << (i - 1));
        }
    }
    bool ring_touches_solid = false;
    if (!solid && !(x > 0 && x < P.nx - 1 && y > 0 && y < P.ny - 1)) {
        // a ring cell never bounces (nb stays 0) but may pull from a solid neighbour, possibly across the
        // periodic wrap: keep its warp out of the pure path, whose in-place (AA) variant assumes that no
        // cell within one step of a pure warp is solid
#pragma unroll
        for (int i = 1; i < 9; i++) {
            int xx = x + dir_ex(i);
            if (xx < 0) xx = P.nx - 1; else if (xx >= P.nx) xx = 0;
            const int mm = row[(ptrdiff_t)dir_ey(i) * P.nx + xx].material;
            ring_touches_solid |= (mm == 2 || mm == 4);
        }
    }
    uint8_t c;
    if (solid) c = CLS_SOLID;
    else if (m == 3 || m == 6) c = CLS_ACCEL;
    else c = (nb || ring_touches_solid) ? CLS_FLUID_NB : CLS_FLUID;
    const size_t cl = (size_t)l * P.pitch + x;
    P.cls[cl] = c;
    P.nbr[cl] = nb;
}

// ------------------------------------------------------------------ class rows of the two halo rows
// Rows y0-1 and y0+h as the two-update kernel sees them: only "solid / inlet-or-force / other" matters there
// (update 1 of a neighbour row never consults the neighbour bits).  Derived from the info halo rows.
__global__ void __launch_bounds__(256) k_derive_halo(const __grid_constant__ SlabParams P, uint8_t *up, uint8_t *dn) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= P.pitch) return;
    uint8_t cu = CLS_SOLID, cd = CLS_SOLID; // pitch padding is never read
    if (x < P.nx) {
        const int mu = P.info[x].material, md = P.info[(size_t)(P.h + 1) * P.nx + x].material;
        cu = (mu == 2 || mu == 4) ? CLS_SOLID : ((mu == 3 || mu == 6) ? CLS_ACCEL : CLS_FLUID_NB);
        cd = (md == 2 || md == 4) ? CLS_SOLID : ((md == 3 || md == 6) ? CLS_ACCEL : CLS_FLUID_NB);
    }
    up[x] = cu;
    dn[x] = cd;
}

// ------------------------------------------------------------------ armed force cells in the info halo rows
// The halo rows are copies of the neighbour slabs' edge rows.  The owners retire an armed inlet / force cell
// (block_iter > 0) on the device — init.wgsl:51-59 at reset (material 1, block_iter 0, velocity 0: zero_v = 1), or
// collide_stream.wgsl:55-62 when its countdown ends (material 1, block_iter 0, velocity kept) — and the copies must
// follow, or the edge row blocks of a sweep would go on forcing a cell that is bulk by now.
__global__ void __launch_bounds__(256) k_retire_halo(const __grid_constant__ SlabParams P, int zero_v) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= P.nx) return;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        LatticeInfo *ip = P.info + (k ? (size_t)(P.h + 1) * P.nx : (size_t)0) + x;
        LatticeInfo in = *ip;
        if ((in.material == 3 || in.material == 6) && in.block_iter > 0) {
            in.material = 1; in.block_iter = 0;
            if (zero_v) { in.vx = 0.0f; in.vy = 0.0f; }
            *ip = in;
        }
    }
}

// ------------------------------------------------------------------ stale values a ring cell would pull
// A fluid cell on the outer ring never receives a bounce-back (boundary.wgsl:19), so when it pulls from a solid
// neighbour it reads whatever that slot holds: 0 after init.wgsl, but a leftover value when the solid was painted
// over live fluid (add_obstacle mid-run) or restored from a checkpoint.  The two-update kernel assumes 0; this
// kernel raises *flag when either buffer holds anything else in such a slot, and the host then keeps to single updates.
__global__ void __launch_bounds__(256) k_ring_check(const __grid_constant__ SlabParams P, unsigned int *flag) {
    const long long per_row = 2;                       // x = 0 and x = nx-1 of every owned row
    const long long n_cols = per_row * P.h;
    const bool top = P.y0 == 0, bot = P.y0 + P.h == P.ny;
    const long long n = n_cols + (top ? P.nx : 0) + (bot ? P.nx : 0);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        int x, l;
        if (t < n_cols) { l = (int)(t >> 1); x = (t & 1) ? P.nx - 1 : 0; }
        else if (top && t < n_cols + P.nx) { l = 0; x = (int)(t - n_cols); }
        else { l = P.h - 1; x = (int)(t - n_cols - (top ? P.nx : 0)); }
        const int mat = P.info[(size_t)(l + 1) * P.nx + x].material;
        if (mat == 2 || mat == 4) continue;
#pragma unroll
        for (int j = 1; j < 9; j++) {
            int xs = x - dir_ex(j);
            if (xs < 0) xs = P.nx - 1; else if (xs >= P.nx) xs = 0;
            const int ls = l - dir_ey(j);
            const int ms = P.info[(size_t)(ls + 1) * P.nx + xs].material; // halo rows hold the wrapped / neighbour rows
            if (ms != 2 && ms != 4) continue;
            for (int b = 0; b < 2; b++) {
                const RowRef rr = row_ref(P, b, ls);
                if (rr.p[(size_t)j * rr.plane + xs] != 0.0f) atomicOr(flag, 1u);
            }
        }
    }
}

// ------------------------------------------------------------------ preset generators
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ bool in_disc(float px, float py, float r) { // fluid/mod.rs:57-59
    return __fsqrt_rn(fadd(fmul(px, px), fmul(py, py))) <= r;
}

// fluid/lattice.rs:26-98 evaluated per cell; fills halo rows too (r = 0 .. h+1).
__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ SlabParams P, int kind, uint64_t seed,
                                                  float solid_fraction) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.nx || r >= P.h + 2) return;
    int y = P.y0 - 1 + r;
    if (y < 0) y += P.ny;
    if (y >= P.ny) y -= P.ny;
    const int nx = P.nx, ny = P.ny;
    int material = 1;
    float vx = 0.0f;
    if (kind == FIELD_ANIMATION_CUSTOM) {
        if (x == 0 || x == nx - 1 || y == 0 || y == ny - 1) material = 2;
    } else if (kind == FIELD_ANIMATION_LID_DRIVEN_CAVITY) {
        if (x == 0 || x == nx - 1 || y == ny - 1) material = 2;
        else if (y == 0) material = 7;
        else if (y == 1) { material = 6; vx = 0.13f; }
    } else { // Poiseuille frame (also the porous preset)
        if (y == 0 || y == ny - 1) material = 2;
        else if (x == 0 || x == nx - 1) material = 7;
        else if (x == 1) { material = 3; vx = 0.12f; }
        else if (x == nx - 2) material = 5;
        else if (kind == FIELD_ANIMATION_POISEUILLE) {
            const float R = 28.0f;
            const float s0x = fsub(fdiv((float)nx, 7.0f), R), s0y = fdiv((float)ny, 2.0f);
            const float s1x = fdiv((float)nx, 5.0f), s1y = fdiv((float)ny, 4.0f);
            const float s2x = s1x, s2y = fmul((float)ny, 0.75f);
            const float px = (float)x, py = (float)y;
            if (in_disc(fsub(px, s0x), fsub(py, s0y), R) || in_disc(fsub(px, s1x), fsub(py, s1y), R) ||
                in_disc(fsub(px, s2x), fsub(py, s2y), R))
                material = 4;
        } else { // LBM_PRESET_POROUS
            const uint64_t hsh = splitmix64(seed ^ (((uint64_t)(uint32_t)y << 32) | (uint64_t)(uint32_t)x));
            const float u = fdiv((float)(hsh >> 40), 16777216.0f);
            if (u < solid_fraction) material = 4;
        }
    }
    LatticeInfo o;
    o.material = material; o.block_iter = -1; o.vx = vx; o.vy = 0.0f;
    P.info[(size_t)r * nx + x] = o;
}

// f64 sum of a dense array (canonicalised AA state)
__global__ void __launch_bounds__(256) k_sum_dense(const float *p, size_t n, double *out) {
    double s = 0.0;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) s += (double)p[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = ws[threadIdx.x];
        for (int o = 4; o > 0; o >>= 1) s += __shfl_down_sync(0xffu, s, o);
        if (threadIdx.x == 0) atomicAdd(out, s);
    }
}

// ------------------------------------------------------------------ total mass (f64)
// One thread takes 4 consecutive cells of a row: nine 128-bit loads in flight per iteration (rows are 128-byte aligned:
// pitch is a multiple of 32 floats), summed in f64.  A pure read stream of 36 B per cell.
__global__ void __launch_bounds__(256) k_mass(const __grid_constant__ SlabParams P, int b, double *out) {
    double s = 0.0;
    const size_t quads_per_row = ((size_t)P.nx + 3) / 4;
    const size_t n = (size_t)P.h * quads_per_row;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
        const size_t l = c / quads_per_row, x = (c - l * quads_per_row) * 4;
        const float *p = P.f[b] + l * P.pitch + x;
        if (x + 4 <= (size_t)P.nx) {
            float4 v[9];
#pragma unroll
            for (int i = 0; i < 9; i++) v[i] = __ldg(reinterpret_cast<const float4 *>(p + i * P.plane));
#pragma unroll
            for (int i = 0; i < 9; i++) s += ((double)v[i].x + (double)v[i].y) + ((double)v[i].z + (double)v[i].w);
        } else {
            for (size_t k = x; k < (size_t)P.nx; k++)
#pragma unroll
                for (int i = 0; i < 9; i++) s += (double)p[(k - x) + i * P.plane];
        }
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = ws[threadIdx.x];
        for (int o = 4; o > 0; o >>= 1) s += __shfl_down_sync(0xffu, s, o);
        if (threadIdx.x == 0) atomicAdd(out, s);
    }
}

}  // namespace lbm
