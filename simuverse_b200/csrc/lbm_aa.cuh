// lbm_aa.cuh — AA-pattern (in-place, single buffer) variant of the D2Q9 step (LBM_FLAG_AA).
//
// One copy of the distributions instead of the A/B ping-pong pair: half the memory, same 72 B/cell of
// traffic per update.  The state alternates between two layouts:
//
//   N ("natural", after an even number of updates): memory == the reference's current buffer, bit for
//     bit, parked bounce-back values in the solid cells' slots included.
//   S ("shifted", after an odd number): the post-collision f*_j of cell x sits where its consumer
//     will read it with a LOCAL access:  A[x+e_j, inv(j)]  if x+e_j is not solid, else  A[x, j].
//
//   update N->S (k_aa_pull):  pull f_i = A[x-e_i, i] exactly like collide_stream.wgsl:43-48, collide,
//     write f*_j to A[x+e_j, inv(j)] — the very location direction inv(j) was pulled from — or to the
//     own slot A[x, j] when x+e_j is solid.  Every cell writes only locations it read (plus own slots
//     that only a solid neighbour would read), so the update is race-free in place.
//   update S->N (k_aa_local): f_i = A[x, inv(i)] (own slots only; a ring cell next to a solid reads the
//     solid's untouched slot A[x-e_i, i] instead, as the reference's pull would), collide, store with
//     the reference layout incl. the bounce-back scatter.  No neighbour reads at all.
//
// lbm_read_distributions canonicalises an S state to the reference layout (k_aa_canonical), so parity
// tests compare the same bytes as for the A/B kernels.  Restrictions: single slab; the previous buffer
// and the on-demand macro field do not exist (the step's inputs are overwritten) — use
// LBM_FLAG_MACRO_EVERY_STEP to get the field.
#pragma once

#include "lbm_step_vec.cuh"

#ifndef LBM_AA_MIN_CTAS
#define LBM_AA_MIN_CTAS 7
#endif

namespace lbm {

__device__ __forceinline__ int wrap_x(const SlabParams &P, int x) { return x < 0 ? P.nx - 1 : (x >= P.nx ? 0 : x); }
__device__ __forceinline__ int wrap_l(const SlabParams &P, int l) { return l < 0 ? P.h - 1 : (l >= P.h ? 0 : l); }
__device__ __forceinline__ bool cell_solid(const SlabParams &P, int x, int l) {
    return P.cls[(size_t)l * P.pitch + x] == CLS_SOLID;
}
__device__ __forceinline__ bool strictly_interior(const SlabParams &P, int x, int l) {
    const int y = P.y0 + l;
    return x > 0 && x < P.nx - 1 && y > 0 && y < P.ny - 1;
}

// ------------------------------------------------------------------ generic per-cell updates
// N -> S
__device__ __forceinline__ void aa_cell_pull(const SlabParams &P, int x, int l) {
    const size_t cl = (size_t)l * P.pitch + x;
    const uint8_t c = P.cls[cl];
    if (c == CLS_SOLID) {
        if (P.macro16 || P.macro32) store_macro(P, x, l, 0.0f, 0.0f, 0.0f, 0.0f);
        return;
    }
    float *A = P.f[0];
    const size_t pl = P.plane;
    int sx[9], sl[9]; // source cell of direction i = target cell of direction inv(i)
    float f[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        sx[i] = wrap_x(P, x - dir_ex(i));
        sl[i] = wrap_l(P, l - dir_ey(i));
        f[i] = A[(size_t)i * pl + (size_t)sl[i] * P.pitch + sx[i]];
    }
    const uint8_t nb = (c == CLS_FLUID) ? 0 : P.nbr[cl];
    collide_cell<0>(P, c, nb, x, l, f);
    const bool interior = strictly_interior(P, x, l);
    A[cl] = f[0];
#pragma unroll
    for (int j = 1; j < 9; j++) {
        const int k = dir_inv(j); // target cell x+e_j == source cell of direction k
        const bool solid = interior ? ((nb >> (j - 1)) & 1) : cell_solid(P, sx[k], sl[k]);
        if (solid) A[(size_t)j * pl + cl] = f[j];
        else A[(size_t)k * pl + (size_t)sl[k] * P.pitch + sx[k]] = f[j];
    }
}

// S -> N
__device__ __forceinline__ void aa_cell_local(const SlabParams &P, int x, int l) {
    const size_t cl = (size_t)l * P.pitch + x;
    const uint8_t c = P.cls[cl];
    if (c == CLS_SOLID) {
        if (P.macro16 || P.macro32) store_macro(P, x, l, 0.0f, 0.0f, 0.0f, 0.0f);
        return;
    }
    float *A = P.f[0];
    const size_t pl = P.plane;
    const bool interior = strictly_interior(P, x, l);
    float f[9];
    f[0] = A[cl];
#pragma unroll
    for (int i = 1; i < 9; i++) {
        f[i] = A[(size_t)dir_inv(i) * pl + cl];
        if (!interior && c != CLS_FLUID) { // ring cell with a solid somewhere around (k_derive marks it)
            const int zx = wrap_x(P, x - dir_ex(i)), zl = wrap_l(P, l - dir_ey(i));
            if (cell_solid(P, zx, zl)) f[i] = A[(size_t)i * pl + (size_t)zl * P.pitch + zx];
        }
    }
    const uint8_t nb = (c == CLS_FLUID) ? 0 : P.nbr[cl];
    collide_cell<0>(P, c, nb, x, l, f);
    A[cl] = f[0];
#pragma unroll
    for (int i = 1; i < 9; i++) {
        if ((nb >> (i - 1)) & 1) { // strictly interior and cell+e_i solid: park (boundary.wgsl:28-31)
            A[(size_t)dir_inv(i) * pl + (size_t)(l + dir_ey(i)) * P.pitch + (x + dir_ex(i))] = f[i];
            A[(size_t)i * pl + cl] = 0.0f;
        } else {
            A[(size_t)i * pl + cl] = f[i];
        }
    }
}

// ------------------------------------------------------------------ vectorised kernels (pure warps)
// Grid as k_step_vec: blockIdx = row * tiles_x + tile, all rows (single slab: the first and last row wrap
// periodically and take the per-cell path).  A pure warp (k_derive: 128 plain-fluid cells, ring cells that
// touch a solid excluded) has no solid within one cell of its span, so all its traffic is 128-bit.
template <bool MACRO>
__global__ void __launch_bounds__(kVecThreads, LBM_AA_MIN_CTAS) k_aa_local(const __grid_constant__ SlabParams P, int tiles_x) {
    const int l = blockIdx.x / tiles_x;
    const int tile = blockIdx.x - l * tiles_x;
    const int x0 = (tile * kVecThreads + threadIdx.x) * 4;
    const int nx = P.nx;
    if (x0 >= nx) return;
    const size_t off = (size_t)l * P.pitch + x0;
    const uint32_t cw = *reinterpret_cast<const uint32_t *>(P.cls + off);
    if (cw != 0 || x0 + 4 > nx || l == 0 || l == P.h - 1) {
#pragma unroll 1
        for (int c = 0; c < 4; c++)
            if (x0 + c < nx) aa_cell_local(P, x0 + c, l);
        return;
    }
    float *__restrict__ a0 = P.f[0] + off;
    const size_t pl = P.plane;
    float4 v[9];
#pragma unroll
    for (int i = 0; i < 9; i++) v[i] = *reinterpret_cast<const float4 *>(a0 + (size_t)dir_inv(i) * pl); // f_i = A[x, inv(i)]
    float F[4][9];
#pragma unroll
    for (int i = 0; i < 9; i++) { F[0][i] = v[i].x; F[1][i] = v[i].y; F[2][i] = v[i].z; F[3][i] = v[i].w; }
    float mrho[4], mux[4], muy[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        moments(F[c], mrho[c], mux[c], muy[c]);
        collide_plain(P.k, mrho[c], mux[c], muy[c], F[c]);
    }
#pragma unroll
    for (int i = 0; i < 9; i++) stg4(a0 + (size_t)i * pl, F[0][i], F[1][i], F[2][i], F[3][i]);
    if (MACRO) {
        const float one[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        if ((nx & 3) == 0) store_macro4<true>(P, (size_t)l * nx + x0, mux, muy, mrho, one);
        else store_macro4<false>(P, (size_t)l * nx + x0, mux, muy, mrho, one);
    }
}

template <bool MACRO>
__global__ void __launch_bounds__(kVecThreads, LBM_AA_MIN_CTAS) k_aa_pull(const __grid_constant__ SlabParams P, int tiles_x) {
    const int l = blockIdx.x / tiles_x;
    const int tile = blockIdx.x - l * tiles_x;
    const int x0 = (tile * kVecThreads + threadIdx.x) * 4;
    const int lane = threadIdx.x & 31;
    const int nx = P.nx;
    const bool in_row = x0 < nx;
    if (l == 0 || l == P.h - 1) { // rows that wrap periodically in y: per-cell path
        if (in_row) {
#pragma unroll 1
            for (int c = 0; c < 4; c++)
                if (x0 + c < nx) aa_cell_pull(P, x0 + c, l);
        }
        return;
    }
    Pulled q = pull_row4(P, 0, l, x0, lane, in_row); // rows l-1 .. l+1 exist
    // A lane is `plain` when its four cells are plain fluid: it then stores 128-bit vectors whose first /
    // last element comes from the neighbouring lane.  Any other lane handles its cells one by one and its
    // neighbours fall back to scalar stores for the element they would have taken from it.  (The location
    // sets of different cells are disjoint, so the order of all these accesses is free.)
    const bool plain = in_row && q.cw == 0 && x0 + 4 <= nx;
    const bool left_plain = __shfl_up_sync(0xffffffffu, plain, 1) && lane > 0;
    const bool right_plain = __shfl_down_sync(0xffffffffu, plain, 1) && lane < 31;
    float F[4][9] = {{q.v0.x, q.v1.x, q.v2.x, q.v3.x, q.v4.x, q.v5.x, q.v6.x, q.v7.x, q.v8.x},
                     {q.v0.y, q.v1.y, q.v2.y, q.v3.y, q.v4.y, q.v5.y, q.v6.y, q.v7.y, q.v8.y},
                     {q.v0.z, q.v1.z, q.v2.z, q.v3.z, q.v4.z, q.v5.z, q.v6.z, q.v7.z, q.v8.z},
                     {q.v0.w, q.v1.w, q.v2.w, q.v3.w, q.v4.w, q.v5.w, q.v6.w, q.v7.w, q.v8.w}};
    float mrho[4], mux[4], muy[4];
#pragma unroll
    for (int c = 0; c < 4; c++) { // non-plain lanes compute throw-away values: no divergence before the shuffles
        moments(F[c], mrho[c], mux[c], muy[c]);
        collide_plain(P.k, mrho[c], mux[c], muy[c], F[c]);
    }
    // f*_j of cell x goes to A[x+e_j, inv(j)]:
    //   j = 1 (1,0) -> plane 3 ; j = 5 (1,-1) -> row l-1, plane 7 ; j = 8 (1,1) -> row l+1, plane 6   (x+1)
    //   j = 3 (-1,0) -> plane 1 ; j = 6 (-1,-1) -> row l-1, plane 8 ; j = 7 (-1,1) -> row l+1, plane 5 (x-1)
    const float l1 = __shfl_up_sync(0xffffffffu, F[3][1], 1), l5 = __shfl_up_sync(0xffffffffu, F[3][5], 1),
                l8 = __shfl_up_sync(0xffffffffu, F[3][8], 1);
    const float r3 = __shfl_down_sync(0xffffffffu, F[0][3], 1), r6 = __shfl_down_sync(0xffffffffu, F[0][6], 1),
                r7 = __shfl_down_sync(0xffffffffu, F[0][7], 1);
    if (!in_row) return;
    if (!plain) {
#pragma unroll 1
        for (int c = 0; c < 4; c++)
            if (x0 + c < nx) aa_cell_pull(P, x0 + c, l);
        return;
    }
    const size_t pl = P.plane;
    float *a0 = P.f[0] + (size_t)l * P.pitch + x0;  // own row
    float *au = a0 - P.pitch, *ad = a0 + P.pitch;    // rows l-1, l+1
    const ptrdiff_t hi = (x0 + 4 >= nx) ? -(ptrdiff_t)x0 : (ptrdiff_t)4;      // column of cell x0+4 (periodic)
    const ptrdiff_t lo = (x0 == 0) ? (ptrdiff_t)(nx - 1) : (ptrdiff_t)-1;      // column of cell x0-1
    stg4(a0, F[0][0], F[1][0], F[2][0], F[3][0]);
    stg4(au + 4 * pl, F[0][2], F[1][2], F[2][2], F[3][2]); // j=2: e=(0,-1) -> row l-1, plane inv(2)=4
    stg4(ad + 2 * pl, F[0][4], F[1][4], F[2][4], F[3][4]); // j=4: e=(0,+1) -> row l+1, plane inv(4)=2
    float *p1 = a0 + 3 * pl, *p5 = au + 7 * pl, *p8 = ad + 6 * pl;
    if (left_plain) {
        stg4(p1, l1, F[0][1], F[1][1], F[2][1]);
        stg4(p5, l5, F[0][5], F[1][5], F[2][5]);
        stg4(p8, l8, F[0][8], F[1][8], F[2][8]);
    } else {
        p1[1] = F[0][1]; p1[2] = F[1][1]; p1[3] = F[2][1];
        p5[1] = F[0][5]; p5[2] = F[1][5]; p5[3] = F[2][5];
        p8[1] = F[0][8]; p8[2] = F[1][8]; p8[3] = F[2][8];
    }
    if (!right_plain) { p1[hi] = F[3][1]; p5[hi] = F[3][5]; p8[hi] = F[3][8]; }
    float *p3 = a0 + 1 * pl, *p6 = au + 8 * pl, *p7 = ad + 5 * pl;
    if (right_plain) {
        stg4(p3, F[1][3], F[2][3], F[3][3], r3);
        stg4(p6, F[1][6], F[2][6], F[3][6], r6);
        stg4(p7, F[1][7], F[2][7], F[3][7], r7);
    } else {
        p3[0] = F[1][3]; p3[1] = F[2][3]; p3[2] = F[3][3];
        p6[0] = F[1][6]; p6[1] = F[2][6]; p6[2] = F[3][6];
        p7[0] = F[1][7]; p7[1] = F[2][7]; p7[2] = F[3][7];
    }
    if (!left_plain) { p3[lo] = F[0][3]; p6[lo] = F[0][6]; p7[lo] = F[0][7]; }
    if (MACRO) {
        const float one[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        if ((nx & 3) == 0) store_macro4<true>(P, (size_t)l * nx + x0, mux, muy, mrho, one);
        else store_macro4<false>(P, (size_t)l * nx + x0, mux, muy, mrho, one);
    }
}

// ------------------------------------------------------------------ canonical view of an S state
// out: dense reference layout [dir][l][x] (9 planes of h*nx).
__global__ void __launch_bounds__(256) k_aa_canonical(const __grid_constant__ SlabParams P, float *out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.nx || l >= P.h) return;
    const float *A = P.f[0];
    const size_t pl = P.plane, n = (size_t)P.h * P.nx;
    const size_t cl = (size_t)l * P.pitch + x, co = (size_t)l * P.nx + x;
    const bool solid = cell_solid(P, x, l);
    const bool interior = strictly_interior(P, x, l);
    out[co] = A[cl];
    for (int j = 1; j < 9; j++) {
        const int yx = wrap_x(P, x + dir_ex(j)), yl = wrap_l(P, l + dir_ey(j)); // the cell this slot points at
        float v;
        if (!solid) {
            if (!cell_solid(P, yx, yl)) v = A[(size_t)dir_inv(j) * pl + (size_t)yl * P.pitch + yx];
            else v = interior ? 0.0f : A[(size_t)j * pl + cl]; // bounced (parked in the solid) / ring: kept
        } else {
            // slot j of a solid is written by the strictly interior fluid cell at cell+e_j (its f*_inv(j))
            const bool writer = !cell_solid(P, yx, yl) && strictly_interior(P, yx, yl);
            v = writer ? A[(size_t)dir_inv(j) * pl + (size_t)yl * P.pitch + yx] : A[(size_t)j * pl + cl];
        }
        out[(size_t)j * n + co] = v;
    }
}

inline cudaError_t launch_step_aa(const SlabParams &P, int parity, cudaStream_t stream) {
    const int tiles_x = (P.nx + kCellsPerCta - 1) / kCellsPerCta;
    const long long blocks = (long long)P.h * tiles_x;
    if (blocks > 2147483647ll) return cudaErrorInvalidConfiguration;
    const bool macro = P.macro16 || P.macro32;
    if (parity == 0) {
        if (macro) k_aa_pull<true><<<(unsigned int)blocks, kVecThreads, 0, stream>>>(P, tiles_x);
        else k_aa_pull<false><<<(unsigned int)blocks, kVecThreads, 0, stream>>>(P, tiles_x);
    } else {
        if (macro) k_aa_local<true><<<(unsigned int)blocks, kVecThreads, 0, stream>>>(P, tiles_x);
        else k_aa_local<false><<<(unsigned int)blocks, kVecThreads, 0, stream>>>(P, tiles_x);
    }
    return cudaGetLastError();
}

}  // namespace lbm
