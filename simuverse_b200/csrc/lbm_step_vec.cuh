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
//   * a 1-byte class plane, read as one 32-bit word per thread, tells whether the thread's four
//     cells are plain interior fluid.  Warps in which every thread says yes are "pure" (vector
//     stores).  All other warps (walls, obstacles, cells that bounce into a solid neighbour,
//     inlet / force cells, the ragged end of a row) are "mixed"; k_scan_mixed counts and lists them
//     whenever the mask changes, and launch_step_vec picks one of three regimes:
//       - mixed warps rare (channel-type masks): ONE launch, k_step_vec<.., INLINE_MIXED=true>; the few
//         non-plain threads take the generic per-cell path of lbm_device.cuh inline;
//       - otherwise k_step_vec<.., false> handles the pure warps and k_step_mixed the listed ones (or
//         every interior warp when most are mixed, e.g. porous media): same vector loads + shuffles,
//         then a warp transpose through shared memory so that the per-cell 32-bit stores (own slots,
//         bounce-back scatter, dead-slot zeros) are issued with consecutive lanes on consecutive
//         cells, i.e. as whole 128-byte lines;
//   * the first and last row of a slab are processed by the CTAs with the lowest block
//     indices through the generic path: their y-neighbours live in the adjacent slab
//     (another GPU's memory mapped over NVLink, or the periodic wrap when there is one
//     slab).  Those CTAs wait on the neighbours' progress flags, do their rows with direct
//     peer loads/stores, and publish this slab's progress — the halo exchange is fused
//     into the step kernel and overlaps the interior rows.
#pragma once

#include <unistd.h>

#include <algorithm>

#include "lbm_device.cuh"

namespace lbm {

inline long getpid_portable() { return (long)getpid(); }

// Progress flags of one slab (uint32 words inside its arena):
//   [0] steps whose edge rows the UP neighbour has finished   (written by that neighbour)
//   [1] same for the DOWN neighbour
//   [2] edge-CTA arrival counter of the running step
//   [3] sticky error word (1 = a wait timed out)
//   [4] steps completed by this slab (device-side, so that launches can be replayed from CUDA graphs)
//   [6..7] u64: nanoseconds edge CTAs spent in wait_neighbours, summed over CTAs;  [8..9] u64: number of such waits
//          (diagnostics for the scaling analysis: lbm_edge_wait_stats)
struct StepSync {
    unsigned int *flags;          // own
    unsigned int *peer_flags[2];  // up / down neighbour's flag words (peer memory)
    int world;
};

// Compile-time tuning knobs (defaults = the measured best; profiles/ has the sweep)
#ifndef LBM_VEC_THREADS
#define LBM_VEC_THREADS 128
#endif
#ifndef LBM_INLINE_MIXED_UNROLL  // unroll of the per-cell loop of non-plain threads (latency vs code size)
#define LBM_INLINE_MIXED_UNROLL 1
#endif
#ifndef LBM_LOAD_HINT   // 0: ld.global.nc (read-only path)  1: + L1::no_allocate  2: plain ld.global
#define LBM_LOAD_HINT 0
#endif
#ifndef LBM_STORE_HINT  // 0: st.global  1: st.global.cs (streaming)  2: st.global.wt
#define LBM_STORE_HINT 0
#endif
constexpr int kInlineMixedUnroll = LBM_INLINE_MIXED_UNROLL;
constexpr int kVecThreads = LBM_VEC_THREADS;     // 4 warps, 512 cells of one row per CTA
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
            const unsigned int step_no = *(volatile unsigned int *)(S.flags + 4); // written by the previous launch
            const unsigned long long t0 = globaltimer_ns();
            while (ld_acquire_sys(S.flags + 0) < step_no || ld_acquire_sys(S.flags + 1) < step_no) {
                if ((long long)(globaltimer_ns() - t0) > kWaitTimeoutNs) {
                    atomicExch(S.flags + 3, 1u);
                    break;
                }
                __nanosleep(200);
            }
            atomicAdd(reinterpret_cast<unsigned long long *>(S.flags + 6), globaltimer_ns() - t0);
            atomicAdd(reinterpret_cast<unsigned long long *>(S.flags + 8), 1ull);
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
                const unsigned int done = *(volatile unsigned int *)(S.flags + 4) + 1u;
                *(volatile unsigned int *)(S.flags + 4) = done; // read by this slab's next launch only
                __threadfence_system();
                st_release_sys(S.peer_flags[0] + 1, done); // I am my up neighbour's DOWN neighbour
                st_release_sys(S.peer_flags[1] + 0, done); // and my down neighbour's UP neighbour
            }
        }
    }
}

__device__ __forceinline__ float4 ldg4(const float *p) {
#if LBM_LOAD_HINT == 1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#elif LBM_LOAD_HINT == 2
    return *reinterpret_cast<const float4 *>(p);
#else
    return __ldg(reinterpret_cast<const float4 *>(p));
#endif
}
__device__ __forceinline__ void stg4(float *p, float a, float b, float c, float d) {
#if LBM_STORE_HINT == 1
    __stcs(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
#elif LBM_STORE_HINT == 2
    __stwt(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
#else
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
#endif
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

// The nine pulled planes of a thread's four cells, x-shifts applied.  All 32 lanes must call this.
struct Pulled {
    float4 v0, v1, v2, v3, v4, v5, v6, v7, v8;
    uint32_t cw;  // class bytes of the four cells (0 when outside the row)
};

__device__ __forceinline__ Pulled pull_row4(const SlabParams &P, int rb, int l, int x0, int lane, bool in_row) {
    Pulled q;
    q.v0 = q.v1 = q.v2 = q.v3 = q.v4 = q.v5 = q.v6 = q.v7 = q.v8 = make_float4(0.f, 0.f, 0.f, 0.f);
    q.cw = 0;
    const int nx = P.nx;
    const size_t off = (size_t)l * P.pitch + x0;
    const float *__restrict__ r0 = P.f[rb] + off;
    const float *__restrict__ ru = r0 - P.pitch; // row y-1: source of e_y = +1 (4,7,8)
    const float *__restrict__ rd = r0 + P.pitch; // row y+1: source of e_y = -1 (2,5,6)
    const size_t pl = P.plane;
    float e1 = 0.f, e5 = 0.f, e8 = 0.f, e3 = 0.f, e6 = 0.f, e7 = 0.f;
    const bool right_edge = in_row && (lane == 31 || x0 + 4 >= nx);
    if (in_row) {
        q.cw = *reinterpret_cast<const uint32_t *>(P.cls + off);
        q.v0 = ldg4(r0);
        q.v1 = ldg4(r0 + 1 * pl);
        q.v3 = ldg4(r0 + 3 * pl);
        q.v2 = ldg4(rd + 2 * pl);
        q.v5 = ldg4(rd + 5 * pl);
        q.v6 = ldg4(rd + 6 * pl);
        q.v4 = ldg4(ru + 4 * pl);
        q.v7 = ldg4(ru + 7 * pl);
        q.v8 = ldg4(ru + 8 * pl);
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
    q.v1 = shift_from_left(q.v1, e1, lane);
    q.v5 = shift_from_left(q.v5, e5, lane);
    q.v8 = shift_from_left(q.v8, e8, lane);
    q.v3 = shift_from_right(q.v3, e3, right_edge);
    q.v6 = shift_from_right(q.v6, e6, right_edge);
    q.v7 = shift_from_right(q.v7, e7, right_edge);
    return q;
}

#define LBM_UNPACK_F(q)                                                                              \
    {{q.v0.x, q.v1.x, q.v2.x, q.v3.x, q.v4.x, q.v5.x, q.v6.x, q.v7.x, q.v8.x},                        \
     {q.v0.y, q.v1.y, q.v2.y, q.v3.y, q.v4.y, q.v5.y, q.v6.y, q.v7.y, q.v8.y},                        \
     {q.v0.z, q.v1.z, q.v2.z, q.v3.z, q.v4.z, q.v5.z, q.v6.z, q.v7.z, q.v8.z},                        \
     {q.v0.w, q.v1.w, q.v2.w, q.v3.w, q.v4.w, q.v5.w, q.v6.w, q.v7.w, q.v8.w}}

// 4 texels of (u.x,u.y,rho,one) as f16 = 32 contiguous bytes
template <bool ALIGNED16>
__device__ __forceinline__ void store_macro4(const SlabParams &P, size_t c, const float (&ux)[4], const float (&uy)[4],
                                             const float (&rho)[4], const float (&one)[4]) {
    if (P.macro16) {
        uint2 t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            __half2 a = __halves2half2(__float2half_rn(ux[k]), __float2half_rn(uy[k]));
            __half2 b = __halves2half2(__float2half_rn(rho[k]), __float2half_rn(one[k]));
            t[k].x = *reinterpret_cast<uint32_t *>(&a);
            t[k].y = *reinterpret_cast<uint32_t *>(&b);
        }
        uint2 *m = reinterpret_cast<uint2 *>(P.macro16) + c;
        if (ALIGNED16) {
            reinterpret_cast<uint4 *>(m)[0] = make_uint4(t[0].x, t[0].y, t[1].x, t[1].y);
            reinterpret_cast<uint4 *>(m)[1] = make_uint4(t[2].x, t[2].y, t[3].x, t[3].y);
        } else {
            m[0] = t[0]; m[1] = t[1]; m[2] = t[2]; m[3] = t[3];
        }
    }
    if (P.macro32) {
        const size_t n = (size_t)P.h * P.nx;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            P.macro32[c + k] = ux[k];
            P.macro32[n + c + k] = uy[k];
            P.macro32[2 * n + c + k] = rho[k];
        }
    }
}

// ------------------------------------------------------------------ pure warps + slab edge rows
// INLINE_MIXED: threads whose four cells are not all plain fluid take the generic per-cell path right
// here (best when mixed warps are rare: no second launch); otherwise mixed warps are skipped and
// k_step_mixed processes them.
template <bool MACRO, bool INLINE_MIXED>
__device__ __forceinline__ void step_vec_block(const SlabParams &P, const StepSync &S, int rb, int tiles_x, unsigned int bid) {
    const int row_k = bid / tiles_x;
    const int tile = bid - row_k * tiles_x;
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
    Pulled q = pull_row4(P, rb, l, x0, lane, in_row);
    if (INLINE_MIXED) {
        if (!in_row) return;
        if (q.cw != 0 || ragged) {
#pragma unroll kInlineMixedUnroll
            for (int c = 0; c < 4; c++)
                if (x0 + c < nx) update_cell<0>(P, rb, x0 + c, l);
            return;
        }
    } else {
        // mixed warps belong to k_step_mixed (same criterion as k_scan_mixed)
        if (__any_sync(0xffffffffu, q.cw != 0 || ragged)) return;
        if (!in_row) return;
    }

    float F[4][9] = LBM_UNPACK_F(q);
    float mrho[4], mux[4], muy[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        moments(F[c], mrho[c], mux[c], muy[c]);
        collide_plain(P.k, mrho[c], mux[c], muy[c], F[c]);
    }
    const size_t pl = P.plane;
    float *__restrict__ w0 = P.f[rb ^ 1] + (size_t)l * P.pitch + x0;
#pragma unroll
    for (int i = 0; i < 9; i++) stg4(w0 + (size_t)i * pl, F[0][i], F[1][i], F[2][i], F[3][i]);
    if (MACRO) {
        const float one[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        if ((nx & 3) == 0) store_macro4<true>(P, (size_t)l * nx + x0, mux, muy, mrho, one);
        else store_macro4<false>(P, (size_t)l * nx + x0, mux, muy, mrho, one);
    }
}

#ifndef LBM_PERSISTENT_CTAS_PER_SM
#define LBM_PERSISTENT_CTAS_PER_SM 0   // 0: one CTA per 512-cell tile;  n: n CTAs per SM striding over the tiles
#endif

template <bool MACRO, bool INLINE_MIXED>
__global__ void __launch_bounds__(kVecThreads) k_step_vec(const __grid_constant__ SlabParams P,
                                                          const __grid_constant__ StepSync S, int rb, int tiles_x,
                                                          unsigned int total_blocks) {
#if LBM_PERSISTENT_CTAS_PER_SM > 0
    for (unsigned int bid = blockIdx.x; bid < total_blocks; bid += gridDim.x)
        step_vec_block<MACRO, INLINE_MIXED>(P, S, rb, tiles_x, bid);
#else
    step_vec_block<MACRO, INLINE_MIXED>(P, S, rb, tiles_x, blockIdx.x);
#endif
}

// ------------------------------------------------------------------ mixed warps
// One warp per entry of the mixed-warp list (or per interior warp when list == nullptr: masks where
// nearly every warp is mixed, e.g. porous media).  Entry id = (l - 1) * warps_per_row + warp_in_row.
//
// Loads are the same 128-bit pulls + shuffles as the pure path.  The nine pulled planes are then
// staged in shared memory (4.5 KB per warp) and re-read transposed, so that lane t owns cells
// t, t+32, t+64, t+96 of the warp's 128-cell span: every per-cell 32-bit store of the warp — own
// slots, the bounce-back scatter of update_cell (lbm_device.cuh), the dead-slot zeros of solid
// cells — covers whole 128-byte lines, and solid lanes simply idle.
template <bool MACRO>
__global__ void __launch_bounds__(kVecThreads) k_step_mixed(const __grid_constant__ SlabParams P, int rb,
                                                            const uint32_t *__restrict__ list, uint32_t count,
                                                            int warps_per_row) {
    const uint32_t slot = blockIdx.x * (kVecThreads / 32) + (threadIdx.x >> 5);
    if (slot >= count) return;
    const uint32_t id = list ? list[slot] : slot;
    const int l = 1 + (int)(id / (uint32_t)warps_per_row);
    const int wx0 = (int)(id % (uint32_t)warps_per_row) * 128;
    const int lane = threadIdx.x & 31;
    const int x0 = wx0 + lane * 4;
    const int nx = P.nx;
    const bool in_row = x0 < nx;
    const bool ragged = in_row && (x0 + 4 > nx);
    Pulled q = pull_row4(P, rb, l, x0, lane, in_row);

    if (__any_sync(0xffffffffu, ragged)) {
        // row end not a multiple of 4: the wrap neighbour of x = nx-1 is not in the registers
        if (in_row) {
#pragma unroll 1
            for (int c = 0; c < 4; c++)
                if (x0 + c < nx) update_cell<0>(P, rb, x0 + c, l);
        }
        return;
    }
    __shared__ float4 stage[kVecThreads / 32][9][32];
    float4(*st)[32] = stage[threadIdx.x >> 5];
    st[0][lane] = q.v0; st[1][lane] = q.v1; st[2][lane] = q.v2;
    st[3][lane] = q.v3; st[4][lane] = q.v4; st[5][lane] = q.v5;
    st[6][lane] = q.v6; st[7][lane] = q.v7; st[8][lane] = q.v8;
    __syncwarp();

    const size_t pl = P.plane;
    const size_t rowoff = (size_t)l * P.pitch;
    float *__restrict__ wrow = P.f[rb ^ 1] + rowoff;
    const int y = P.y0 + l;
#pragma unroll 2
    for (int c = 0; c < 4; c++) {
        const int xc = wx0 + 32 * c + lane;
        if (xc >= nx) continue;
        const uint32_t cc = P.cls[rowoff + xc];
        const uint32_t nb = P.nbr[rowoff + xc];
        float *wc = wrow + xc;
        if (cc == CLS_SOLID) {
            if (MACRO) store_macro(P, xc, l, 0.0f, 0.0f, 0.0f, 0.0f);
            zero_dead_slots(P, wc, nb, xc, y);
            continue;
        }
        if (cc == CLS_ACCEL) { // inlet / force cell: fully generic path (rare)
            update_cell<0>(P, rb, xc, l);
            continue;
        }
        if (cc == CLS_FLIPPED) P.cls[rowoff + xc] = nb ? CLS_FLUID_NB : CLS_FLUID; // plain fluid from now on
        float f[9];
#pragma unroll
        for (int i = 0; i < 9; i++) f[i] = reinterpret_cast<const float *>(st[i])[32 * c + lane];
        float rho, ux, uy;
        moments(f, rho, ux, uy);
        collide_plain(P.k, rho, ux, uy, f);
        if (MACRO) store_macro(P, xc, l, ux, uy, rho, 1.0f);
        wc[0] = f[0];
        if (nb == 0) {
#pragma unroll
            for (int i = 1; i < 9; i++) wc[(size_t)i * pl] = f[i];
        } else {
#pragma unroll
            for (int i = 1; i < 9; i++) {
                const bool bounce = (nb >> (i - 1)) & 1u;
                wc[(size_t)i * pl] = bounce ? 0.0f : f[i];
                if (bounce) wc[(ptrdiff_t)((size_t)dir_inv(i) * pl) + (ptrdiff_t)dir_ey(i) * P.pitch + dir_ex(i)] = f[i];
            }
        }
    }
}

#undef LBM_UNPACK_F

// Lists the interior warps that k_step_vec leaves to k_step_mixed: any cell of the warp's 128-cell
// span is not plain fluid, or the span contains the ragged end of a row.
__global__ void __launch_bounds__(256) k_scan_mixed(const __grid_constant__ SlabParams P, uint32_t *list, uint32_t *count,
                                                    int warps_per_row) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = (uint32_t)warps_per_row * (uint32_t)max(P.h - 2, 0);
    if (id >= total) return;
    const int l = 1 + (int)(id / (uint32_t)warps_per_row);
    const int wx0 = (int)(id % (uint32_t)warps_per_row) * 128;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(P.cls + (size_t)l * P.pitch + wx0);
    bool mixed = false;
    for (int k = 0; k < 32; k++) {
        const int x0 = wx0 + 4 * k;
        if (x0 >= P.nx) break;
        if (x0 + 4 > P.nx || w[k] != 0) { mixed = true; break; }
    }
    if (mixed) list[atomicAdd(count, 1u)] = id;
}

// wait / signal as stand-alone launches around the generic kernel (LBM_FLAG_KERNEL_GENERIC)
__global__ void k_wait(const __grid_constant__ StepSync S) { wait_neighbours(S); }
__global__ void k_signal(const __grid_constant__ StepSync S) { signal_neighbours(S, 1u); }

struct MixedList {
    uint32_t *list = nullptr;   // device: ids of mixed interior warps (capacity = all interior warps)
    uint32_t *count_dev = nullptr;
    uint32_t count = 0;         // host copy
    uint32_t total = 0;         // interior warps
    int warps_per_row = 0;
    bool everywhere = false;    // most warps are mixed: run k_step_mixed over all of them, no list
    bool rare = true;           // few warps are mixed: k_step_vec handles them inline, no second launch
};

// One lattice update of a slab.  Mixed warps rare (<= 1/8 of the interior warps; channel-type masks):
// one launch of k_step_vec<.., INLINE_MIXED>.  Otherwise k_step_vec (edge rows + pure warps) followed by
// k_step_mixed over the list, or over every interior warp when more than half are mixed (porous media).
// The two kernels write disjoint cells and both only read buffer rb, so their order is free.
// Returns the number of kernels launched through *launched.
inline cudaError_t launch_step_vec(const SlabParams &P, const StepSync &S, const MixedList &M, int rb, cudaStream_t stream,
                                   int *launched) {
    const int tiles_x = (P.nx + kCellsPerCta - 1) / kCellsPerCta;
    const bool macro = P.macro16 || P.macro32;
    // block rows: 0 -> row 0, 1 -> row h-1, k >= 2 -> row k-1 (edge rows are dispatched first)
    // (only the two edge rows when k_step_mixed takes every interior warp)
    const long long rows_k = (P.h >= 2 && (M.rare || !M.everywhere)) ? P.h : 2;
    const long long blocks = rows_k * tiles_x;
    if (blocks > 2147483647ll) return cudaErrorInvalidConfiguration;
    *launched = 1;
    unsigned int grid = (unsigned int)blocks;
#if LBM_PERSISTENT_CTAS_PER_SM > 0
    grid = (unsigned int)std::min<long long>(blocks, 148ll * LBM_PERSISTENT_CTAS_PER_SM);
#endif
    const unsigned int total = (unsigned int)blocks;
    if (M.rare) {
        if (macro) k_step_vec<true, true><<<grid, kVecThreads, 0, stream>>>(P, S, rb, tiles_x, total);
        else k_step_vec<false, true><<<grid, kVecThreads, 0, stream>>>(P, S, rb, tiles_x, total);
        return cudaGetLastError();
    }
    if (macro) k_step_vec<true, false><<<grid, kVecThreads, 0, stream>>>(P, S, rb, tiles_x, total);
    else k_step_vec<false, false><<<grid, kVecThreads, 0, stream>>>(P, S, rb, tiles_x, total);
    const uint32_t n = M.everywhere ? M.total : M.count;
    if (n > 0) {
        const uint32_t *list = M.everywhere ? nullptr : M.list;
        const unsigned int grid = (n + kVecThreads / 32 - 1) / (kVecThreads / 32);
        if (macro) k_step_mixed<true><<<grid, kVecThreads, 0, stream>>>(P, rb, list, n, M.warps_per_row);
        else k_step_mixed<false><<<grid, kVecThreads, 0, stream>>>(P, rb, list, n, M.warps_per_row);
        *launched = 2;
    }
    return cudaGetLastError();
}

}  // namespace lbm
