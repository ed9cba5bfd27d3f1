// lbm_step_vec.cuh — the production D2Q9 step: one fused pull / collide / bounce-back kernel.
//
// Mapping (HBM-bound stencil, 72 B/cell algorithmic: 9 f32 in + 9 f32 out; no tensor cores —
// nothing here is a contraction):
//   * one thread owns 4 consecutive cells of a row: every plane is read and written with
//     128-bit, fully coalesced accesses (a warp moves 512 contiguous bytes per plane);
//   * the +-1 x shifts of the six directions with e_x != 0 come from the neighbouring lane
//     by warp shuffle; only lane 0 / lane 31 issue one extra scalar load (periodic wrap
//     folded into its address, layout_and_fn.wgsl:40-44);
//   * y-neighbour rows are distinct data (every f_i is pulled by exactly one cell), so
//     rows are streamed straight into registers — no shared-memory staging is needed for
//     reuse; the only on-chip exchange is the shuffle;
//   * cells that are not plain interior fluid (walls, obstacles, inlet / force cells, cells
//     that bounce into a solid neighbour, the ragged end of a row) take the generic
//     per-cell path of lbm_device.cuh; a 1-byte class plane decides, read as one 32-bit
//     word per thread;
//   * the first and last row of a slab are processed by the CTAs with the lowest block
//     indices through the generic path: their y-neighbours live in the adjacent slab
//     (another GPU's memory mapped over NVLink, or the periodic wrap when there is one
//     slab).  Those CTAs wait on the neighbours' progress flags, do their rows with direct
//     peer loads/stores, and publish this slab's progress — the halo exchange is fused
//     into the step kernel and overlaps the interior rows.
#pragma once

#include <unistd.h>

#include "lbm_device.cuh"

namespace lbm {

inline long getpid_portable() { return (long)getpid(); }

// Progress flags of one slab (uint32 words inside its arena):
//   [0] steps whose edge rows the UP neighbour has finished   (written by that neighbour)
//   [1] same for the DOWN neighbour
//   [2] edge-CTA arrival counter of the running step
//   [3] sticky error word (1 = a wait timed out)
struct StepSync {
    unsigned int *flags;          // own
    unsigned int *peer_flags[2];  // up / down neighbour's flag words (peer memory)
    int world;
    unsigned int step_no;         // steps completed by this slab before the current one
};

constexpr int kVecThreads = 128;                 // 4 warps, 512 cells of one row per CTA
constexpr int kCellsPerCta = kVecThreads * 4;
constexpr long long kWaitTimeoutNs = 4000000000ll;  // 4 s: never hang the GPU on a lost neighbour

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Edge rows may start once both neighbours have finished the edge rows of the previous step:
// then the rows this slab pulls from are final, and nobody still reads the rows it overwrites.
__device__ __forceinline__ void wait_neighbours(const StepSync &S) {
    if (S.world > 1) {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = globaltimer_ns();
            while (ld_acquire_sys(S.flags + 0) < S.step_no || ld_acquire_sys(S.flags + 1) < S.step_no) {
                if ((long long)(globaltimer_ns() - t0) > kWaitTimeoutNs) {
                    atomicExch(S.flags + 3, 1u);
                    break;
                }
                __nanosleep(200);
            }
        }
        __syncthreads();
    }
}

// Called by every thread of every edge CTA after its stores; the last CTA publishes.
__device__ __forceinline__ void signal_neighbours(const StepSync &S, unsigned int n_edge_ctas) {
    if (S.world > 1) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int arrived = atomicAdd(S.flags + 2, 1u) + 1u;
            if (arrived == n_edge_ctas) {
                atomicExch(S.flags + 2, 0u);
                __threadfence_system();
                st_release_sys(S.peer_flags[0] + 1, S.step_no + 1u); // I am my up neighbour's DOWN neighbour
                st_release_sys(S.peer_flags[1] + 0, S.step_no + 1u); // and my down neighbour's UP neighbour
            }
        }
    }
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void stg4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}

// Shift towards +x: cell x takes the value of x-1.  `edge` is the element left of the lane's
// vector, only meaningful in lane 0.
__device__ __forceinline__ float4 shift_from_left(float4 v, float edge, int lane) {
    float l = __shfl_up_sync(0xffffffffu, v.w, 1);
    if (lane == 0) l = edge;
    return make_float4(l, v.x, v.y, v.z);
}
// Shift towards -x: cell x takes the value of x+1.
__device__ __forceinline__ float4 shift_from_right(float4 v, float edge, bool use_edge) {
    float r = __shfl_down_sync(0xffffffffu, v.x, 1);
    if (use_edge) r = edge;
    return make_float4(v.y, v.z, v.w, r);
}

#define LBM_CELL(c, comp)                                                                  \
    {                                                                                      \
        float f[9] = {v0.comp, v1.comp, v2.comp, v3.comp, v4.comp, v5.comp, v6.comp, v7.comp, v8.comp}; \
        float rho, ux, uy;                                                                 \
        moments(f, rho, ux, uy);                                                           \
        collide_plain(P.k, rho, ux, uy, f);                                                \
        v0.comp = f[0]; v1.comp = f[1]; v2.comp = f[2]; v3.comp = f[3]; v4.comp = f[4];    \
        v5.comp = f[5]; v6.comp = f[6]; v7.comp = f[7]; v8.comp = f[8];                    \
        mrho[c] = rho; mux[c] = ux; muy[c] = uy;                                           \
    }

template <bool MACRO>
__global__ void __launch_bounds__(kVecThreads) k_step_vec(const __grid_constant__ SlabParams P,
                                                          const __grid_constant__ StepSync S, int rb, int tiles_x) {
    const int row_k = blockIdx.x / tiles_x;
    const int tile = blockIdx.x - row_k * tiles_x;
    const int x0 = (tile * kVecThreads + threadIdx.x) * 4;

    if (row_k < 2) {
        // ---- slab edge rows: generic path over neighbour-slab memory, fused halo exchange
        if (row_k == 1 && P.h < 2) return; // cannot happen for world > 1 (create enforces h >= 2)
        const int l = (row_k == 0) ? 0 : P.h - 1;
        wait_neighbours(S);
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (x0 + c < P.nx) update_cell<0>(P, rb, x0 + c, l);
        signal_neighbours(S, 2u * (unsigned int)tiles_x);
        return;
    }
    const int l = row_k - 1; // interior rows 1 .. h-2
    const int lane = threadIdx.x & 31;
    const int nx = P.nx;
    const bool in_row = x0 < nx;
    const bool ragged = in_row && (x0 + 4 > nx);

    const size_t off = (size_t)l * P.pitch + x0;
    const float *__restrict__ r0 = P.f[rb] + off;
    const float *__restrict__ ru = r0 - P.pitch; // row y-1: source of e_y = +1 (4,7,8)
    const float *__restrict__ rd = r0 + P.pitch; // row y+1: source of e_y = -1 (2,5,6)
    const size_t pl = P.plane;

    float4 v0, v1, v2, v3, v4, v5, v6, v7, v8;
    v0 = v1 = v2 = v3 = v4 = v5 = v6 = v7 = v8 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cw = 0;
    float e1 = 0.f, e5 = 0.f, e8 = 0.f, e3 = 0.f, e6 = 0.f, e7 = 0.f;
    const bool right_edge = in_row && (lane == 31 || x0 + 4 >= nx);
    if (in_row) {
        cw = *reinterpret_cast<const uint32_t *>(P.cls + off);
        v0 = ldg4(r0);
        v1 = ldg4(r0 + 1 * pl);
        v3 = ldg4(r0 + 3 * pl);
        v2 = ldg4(rd + 2 * pl);
        v5 = ldg4(rd + 5 * pl);
        v6 = ldg4(rd + 6 * pl);
        v4 = ldg4(ru + 4 * pl);
        v7 = ldg4(ru + 7 * pl);
        v8 = ldg4(ru + 8 * pl);
        if (lane == 0) {
            const ptrdiff_t dl = (x0 == 0) ? (ptrdiff_t)(nx - 1) : (ptrdiff_t)-1; // periodic wrap
            e1 = __ldg(r0 + 1 * pl + dl);
            e5 = __ldg(rd + 5 * pl + dl);
            e8 = __ldg(ru + 8 * pl + dl);
        }
        if (right_edge) {
            const ptrdiff_t dr = (x0 + 4 >= nx) ? -(ptrdiff_t)x0 : (ptrdiff_t)4; // wrap to column 0
            e3 = __ldg(r0 + 3 * pl + dr);
            e6 = __ldg(rd + 6 * pl + dr);
            e7 = __ldg(ru + 7 * pl + dr);
        }
    }
    // all 32 lanes take part in the shuffles, active or not
    v1 = shift_from_left(v1, e1, lane);
    v5 = shift_from_left(v5, e5, lane);
    v8 = shift_from_left(v8, e8, lane);
    v3 = shift_from_right(v3, e3, right_edge);
    v6 = shift_from_right(v6, e6, right_edge);
    v7 = shift_from_right(v7, e7, right_edge);
    if (!in_row) return;

    if (cw != 0 || ragged) {
        // not four plain fluid cells: generic per-cell path (re-reads hit L1/L2)
#pragma unroll 1
        for (int c = 0; c < 4; c++)
            if (x0 + c < nx) update_cell<0>(P, rb, x0 + c, l);
        return;
    }

    float mrho[4], mux[4], muy[4];
    LBM_CELL(0, x)
    LBM_CELL(1, y)
    LBM_CELL(2, z)
    LBM_CELL(3, w)

    float *__restrict__ w0 = P.f[rb ^ 1] + off;
    *reinterpret_cast<float4 *>(w0) = v0;
    *reinterpret_cast<float4 *>(w0 + 1 * pl) = v1;
    *reinterpret_cast<float4 *>(w0 + 2 * pl) = v2;
    *reinterpret_cast<float4 *>(w0 + 3 * pl) = v3;
    *reinterpret_cast<float4 *>(w0 + 4 * pl) = v4;
    *reinterpret_cast<float4 *>(w0 + 5 * pl) = v5;
    *reinterpret_cast<float4 *>(w0 + 6 * pl) = v6;
    *reinterpret_cast<float4 *>(w0 + 7 * pl) = v7;
    *reinterpret_cast<float4 *>(w0 + 8 * pl) = v8;

    if (MACRO) {
        const size_t c = (size_t)l * nx + x0;
        if (P.macro16) {
            // 4 texels of (u.x,u.y,rho,1) f16 = 32 bytes; c*8 is 16-byte aligned when nx%4==0,
            // otherwise fall back to 8-byte stores
            uint2 t[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                __half2 a = __halves2half2(__float2half_rn(mux[k]), __float2half_rn(muy[k]));
                __half2 b = __halves2half2(__float2half_rn(mrho[k]), __float2half_rn(1.0f));
                t[k].x = *reinterpret_cast<uint32_t *>(&a);
                t[k].y = *reinterpret_cast<uint32_t *>(&b);
            }
            uint2 *m = reinterpret_cast<uint2 *>(P.macro16) + c;
            if ((nx & 3) == 0) {
                reinterpret_cast<uint4 *>(m)[0] = make_uint4(t[0].x, t[0].y, t[1].x, t[1].y);
                reinterpret_cast<uint4 *>(m)[1] = make_uint4(t[2].x, t[2].y, t[3].x, t[3].y);
            } else {
                m[0] = t[0]; m[1] = t[1]; m[2] = t[2]; m[3] = t[3];
            }
        }
        if (P.macro32) {
            const size_t n = (size_t)P.h * nx;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                P.macro32[c + k] = mux[k];
                P.macro32[n + c + k] = muy[k];
                P.macro32[2 * n + c + k] = mrho[k];
            }
        }
    }
}

#undef LBM_CELL

// wait / signal as stand-alone launches around the generic kernel (LBM_FLAG_KERNEL_GENERIC)
__global__ void k_wait(const __grid_constant__ StepSync S) { wait_neighbours(S); }
__global__ void k_signal(const __grid_constant__ StepSync S) { signal_neighbours(S, 1u); }

inline cudaError_t launch_step_vec(const SlabParams &P, const StepSync &S, int rb, cudaStream_t stream) {
    const int tiles_x = (P.nx + kCellsPerCta - 1) / kCellsPerCta;
    // block rows: 0 -> row 0, 1 -> row h-1, k >= 2 -> row k-1 (edge rows are dispatched first)
    const long long rows_k = (P.h >= 2) ? P.h : 2;
    const long long blocks = rows_k * tiles_x;
    if (blocks > 2147483647ll) return cudaErrorInvalidConfiguration;
    if (P.macro16 || P.macro32)
        k_step_vec<true><<<(unsigned int)blocks, kVecThreads, 0, stream>>>(P, S, rb, tiles_x);
    else
        k_step_vec<false><<<(unsigned int)blocks, kVecThreads, 0, stream>>>(P, S, rb, tiles_x);
    return cudaGetLastError();
}

}  // namespace lbm
