// lbm_b200.cu — C ABI (include/lbm_b200.h) over the CUDA kernels: handle, memory layout,
// launch logic.  There is deliberately no CPU code path for any compute entry point: with no
// CUDA device lbm_create fails with LBM_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"
#include "lbm_kernels.cuh"
#include "lbm_particles.cuh"
#include "lbm_step_vec.cuh"
#include "lbm_aa.cuh"
#include "lbm_fused.cuh"

using namespace lbm;

namespace {

std::string g_create_error;

// What a slab publishes to its neighbours (fits LbmIpcBlob).
struct IpcPayload {
    uint32_t magic;
    int32_t pid;
    int32_t device;
    int32_t rank, world;
    int32_t nx, ny, y0, h, pitch;
    uint64_t plane;           // floats
    uint64_t f_off[2];        // byte offsets of f[0], f[1] in the arena
    uint64_t spare_off;       // byte offset of the third buffer (multi-slab handles; see materialize_prev)
    uint64_t flag_off;        // byte offset of the progress flags
    uint64_t arena_bytes;
    uint64_t local_ptr;       // arena address in the exporting process (same-process attach)
    cudaIpcMemHandle_t handle;
};
static_assert(sizeof(IpcPayload) <= sizeof(LbmIpcBlob), "blob too small");
constexpr uint32_t kIpcMagic = 0x4c424d31u; // "LBM1"

}  // namespace

struct LbmSim {
    LbmDesc d{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    SlabParams P{};
    char *arena = nullptr;      // f[0], f[1], sync flags (one allocation, one IPC handle)
    size_t arena_bytes = 0;
    size_t f_off[2] = {0, 0};
    size_t flag_off = 0;
    double *d_mass = nullptr;
    float *scratch32 = nullptr; // 3 f32 planes for the on-demand macro read
    __half *scratch16 = nullptr; // RGBA16F texels for the on-demand macro read
    __half *curl16 = nullptr;    // RGBA16F texels of lbm_read_curl
    float4 *present32 = nullptr; // fragment outputs of lbm_read_present (present_cap pixels)
    size_t present_cap = 0;
    // second macro texture + copy stream for lbm_read_macro_async (pipelined field read-back)
    __half *macro_buf[2] = {nullptr, nullptr};
    int macro_cur = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_copied[2] = {nullptr, nullptr};
    bool copy_pending[2] = {false, false};
    uint64_t macro_writes = 0, macro_writes_at_flip = ~0ull; // to serve the newest texture after a flip
    float *scratch_dense = nullptr; // 9 dense planes: canonical view of an AA state in its shifted layout
    bool aa = false;
    uint64_t steps_since_reset = 0;
    bool have_uniform = false;
    bool have_info = false;
    LbmUniform u{};
    FieldUniform field{};
    ParticleUniform pu{};
    bool have_pu = false;
    TrajectoryParticle *particles = nullptr;
    uint64_t n_particles = 0;
    Pixel *canvas = nullptr;
    int canvas_w = 0, canvas_h = 0;
    int swap = 0;
    uint64_t launches = 0;
    // neighbours (multi-slab)
    void *peer_base[2] = {nullptr, nullptr}; // IPC-opened arenas (up, down); nullptr when same-process
    bool peer_ipc[2] = {false, false};
    StepSync sync{};
    MixedList mixed{};
    bool attached = false;
    // CUDA graphs of kGraphSteps consecutive steps / one frame (single-slab handles only): removes the
    // per-launch host cost and most of the inter-kernel gap; rebuilt whenever a kernel parameter changes
    cudaGraphExec_t graph_steps[2] = {nullptr, nullptr}; // indexed by the swap index of the first step
    cudaGraphExec_t graph_frame = nullptr;
    uint64_t graph_steps_kernels[2] = {0, 0}; // kernels inside each captured graph (gpu_launches accounting)
    uint64_t graph_frame_kernels = 0;
    // two updates per sweep (lbm_fused.cuh)
    FuseGeom fuse{};
    int2 *d_fuse_rows = nullptr;           // FuseGeom::items
    unsigned int *d_fuse_flags = nullptr;  // [0] largest armed block_iter seen by k_derive, [1] k_ring_check verdict
    uint8_t *cls_halo = nullptr;           // multi-slab: class rows y0-1 and y0+h (2 * pitch bytes)
    int64_t countdown_left = 0;            // upper bound of the updates during which a force cell may still count down / retire
    bool fuse_blocked = false;             // a ring cell would pull a stale value out of a solid (see k_ring_check)
    bool ring_check_needed = false;
    bool mask_written_since_reset = false;
    bool use_masked = true;                // sweep kernel instance with the inline masked path (LBM_FUSE_MASKED=0: without)
    bool mixed_dirty = false;              // the mask changed: the mixed-warp list is rebuilt before the next single update
    bool halo_retire_pending = false;      // multi-slab: armed force cells in the info halo rows retire when the countdown ends
    // pinned staging ring of lbm_write_lattice_info: the caller's bytes are copied here and travel asynchronously, so
    // a mask write never waits for the device (and the caller may reuse its buffer at once)
    static constexpr int kStageSlots = 4;
    static constexpr size_t kStageBytes = 4u << 20;
    char *stage[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool stage_used[kStageSlots] = {false, false, false, false};
    int stage_next = 0;
    size_t stage_cursor = 0;
    float *prev_spare = nullptr;           // single slab: persistent spare buffer of materialize_prev (allocated on first use)
    float *prev_spare_alloc = nullptr;     // ... the allocation itself: after a pointer exchange the spare is a part of the arena
    bool spare_keep_dirty = true;          // the slots no update writes may differ between the spare and the real buffers
    // multi-slab only: a third distribution buffer in the arena (peer-visible), so that the buffer a sweep leaves two
    // updates behind can be recomputed by an ordinary, neighbour-synchronised update and swapped in
    float *spare_f = nullptr, *spare_up = nullptr, *spare_dn = nullptr;
    size_t spare_off = 0;
    bool prev_stale = false;               // after a two-update sweep the non-current buffer holds t, not t+1
    int flip = 0;                          // parity of the buffer-pointer exchanges done by two-update sweeps
    cudaGraphExec_t graph_pairs[4] = {nullptr, nullptr, nullptr, nullptr}; // [flip * 2 + swap]: kGraphSteps / 2 sweeps
    uint64_t graph_pairs_kernels[4] = {0, 0, 0, 0};
    // frames with tracer particles as sweeps: the texture update 1 of a sweep stores into; graphs of 4 / 8 frames
    // (sweep, particles(t+1), particles(t+2)) — an even number of sweeps leaves the buffer pointers unchanged
    __half *macro_mid = nullptr;
    // Frames with tracer particles, overlapped: the two particle passes of frame f only read the textures sweep f stored
    // and sweep f+1 only reads the distributions, so inside a captured run of frames the passes run on a second stream
    // beside the next sweep.  Two texture sets alternate (set 0 = the handle's own macro_mid / macro16, set 1 = these).
    __half *macro_mid2 = nullptr, *macro_alt = nullptr;
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_sweep = nullptr, ev_part[2] = {nullptr, nullptr};
    bool overlap_particles = true;         // LBM_PARTICLE_OVERLAP=0: everything on one stream (A/B runs)
    bool pdl = false;                      // LBM_PDL=1: sweeps are launched with programmatic stream serialization
    cudaGraphExec_t graph_pframes[4] = {nullptr, nullptr, nullptr, nullptr}; // [flip * 2 + (8-frame run ? 1 : 0)]
    uint64_t graph_pframes_kernels[4] = {0, 0, 0, 0};
    uint64_t fused_sweeps = 0;
    std::string err;
};

namespace {

int fail(LbmSim *s, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (s) s->err = buf; else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(s, e__ == cudaErrorMemoryAllocation ? LBM_ERR_OUT_OF_MEMORY : LBM_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int check_launch(LbmSim *s, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(s, LBM_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    s->launches++;
    return LBM_OK;
}

dim3 grid2d(int nx, int rows, dim3 block) { return dim3((nx + block.x - 1) / block.x, (rows + block.y - 1) / block.y, 1); }

void invalidate_graphs(LbmSim *s);
void invalidate_step_graphs(LbmSim *s);

// Re-derives the class / neighbour planes of owned rows [l0, l1) from the info buffer.  Asynchronous: nothing is read
// back here.  The mixed-warp list of the single-update kernels is rebuilt lazily (ensure_mixed), and only the graphs
// that contain those kernels are dropped — a two-update sweep does not depend on the mask geometry.
// want_armed: also collect the largest armed block_iter of the rows into d_fuse_flags[0] (read by the caller).
int derive_rows(LbmSim *s, int l0, int l1, bool want_armed = false) {
    l0 = std::max(l0, 0);
    l1 = std::min(l1, s->P.h);
    if (l0 >= l1) return LBM_OK;
    dim3 block(64, 4);
    if (want_armed) CU(cudaMemsetAsync(s->d_fuse_flags, 0, sizeof(unsigned int), s->stream));
    k_derive<<<grid2d(s->P.nx, l1 - l0, block), block, 0, s->stream>>>(s->P, l0, l1, want_armed ? s->d_fuse_flags : nullptr);
    int rc = check_launch(s, "k_derive");
    if (rc) return rc;
    if (s->cls_halo) {
        k_derive_halo<<<(s->P.pitch + 255) / 256, 256, 0, s->stream>>>(s->P, s->cls_halo, s->cls_halo + s->P.pitch);
        if ((rc = check_launch(s, "k_derive_halo"))) return rc;
    }
    s->mixed_dirty = true;
    invalidate_step_graphs(s); // launch geometry of the single-update kernels is baked into those graphs
    return LBM_OK;
}

// The list of warps k_step_vec leaves to k_step_mixed, rebuilt after a mask change — before the next single update, not
// at the mask write (one full pass over the class plane and a device -> host count).
int ensure_mixed(LbmSim *s) {
    if (!s->mixed_dirty) return LBM_OK;
    s->mixed_dirty = false;
    MixedList &M = s->mixed;
    if (M.total == 0 || s->aa || (s->d.flags & LBM_FLAG_KERNEL_GENERIC)) return LBM_OK;
    CU(cudaMemsetAsync(M.count_dev, 0, sizeof(uint32_t), s->stream));
    k_scan_mixed<<<(M.total + 255) / 256, 256, 0, s->stream>>>(s->P, M.list, M.count_dev, M.warps_per_row);
    int rc = check_launch(s, "k_scan_mixed");
    if (rc) return rc;
    CU(cudaMemcpyAsync(&M.count, M.count_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    M.everywhere = M.count > M.total / 2;
    // few mixed warps, or a lattice so small that a second launch costs more than a few slow threads
    M.rare = M.count <= M.total / 8 || M.total <= 8192;
    return LBM_OK;
}

bool uniform_is_d2q9(const LbmUniform *u) {
    static const float ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    static const float ey[9] = {0, 0, -1, 0, 1, -1, -1, 1, 1};
    static const int inv[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    for (int i = 0; i < 9; i++) {
        if (u->e_w_max[i][0] != ex[i] || u->e_w_max[i][1] != ey[i]) return false;
        if (u->inversed_direction[i][0] != inv[i]) return false;
    }
    return true;
}

int ensure_scratch32(LbmSim *s) {
    if (s->scratch32) return LBM_OK;
    CU(cudaMalloc(&s->scratch32, sizeof(float) * 3 * (size_t)s->P.h * s->P.nx));
    return LBM_OK;
}

int ensure_scratch16(LbmSim *s) {
    if (s->scratch16) return LBM_OK;
    CU(cudaMalloc(&s->scratch16, sizeof(__half) * 4 * (size_t)s->P.h * s->P.nx));
    return LBM_OK;
}

// init.wgsl:62 leaves the macro texture at (0,0,0,1) everywhere
__global__ void k_macro_after_init(const __grid_constant__ SlabParams P) {
    const size_t n = (size_t)P.h * P.nx;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
        store_macro(P, (int)(c % P.nx), (int)(c / P.nx), 0.0f, 0.0f, 0.0f, 1.0f);
}

int launch_step(LbmSim *s, int rb) {
    int rc = LBM_OK;
    if (s->aa) {
        cudaError_t e = launch_step_aa(s->P, rb, s->stream); // rb = parity: 0 pulls (N->S), 1 is local (S->N)
        if (e != cudaSuccess) return fail(s, LBM_ERR_CUDA, "launch of k_aa_* failed: %s", cudaGetErrorString(e));
        s->launches++;
    } else if (s->d.flags & LBM_FLAG_KERNEL_GENERIC) {
        const bool multi = s->d.world > 1;
        if (multi) {
            k_wait<<<1, 32, 0, s->stream>>>(s->sync);
            if ((rc = check_launch(s, "k_wait"))) return rc;
        }
        dim3 block(64, 4);
        k_step_generic<0><<<grid2d(s->P.nx, s->P.h, block), block, 0, s->stream>>>(s->P, rb, 0, s->P.h);
        if ((rc = check_launch(s, "k_step_generic"))) return rc;
        if (multi) {
            k_signal<<<1, 32, 0, s->stream>>>(s->sync);
            if ((rc = check_launch(s, "k_signal"))) return rc;
        }
    } else {
        int n = 0;
        cudaError_t e = launch_step_vec(s->P, s->sync, s->mixed, rb, s->stream, &n);
        if (e != cudaSuccess) return fail(s, LBM_ERR_CUDA, "launch of k_step_vec / k_step_mixed failed: %s", cudaGetErrorString(e));
        s->launches += n;
    }
    s->steps_since_reset++;
    s->macro_writes++;
    s->prev_stale = false; // the buffer just written is the true t+1 again
    if (s->countdown_left > 0) s->countdown_left--;
    return LBM_OK;
}

int ready_to_step(LbmSim *s) {
    if (!s->have_uniform) return fail(s, LBM_ERR_STATE, "lbm_write_uniform has not been called");
    if (!s->have_info) return fail(s, LBM_ERR_STATE, "no lattice info uploaded or generated");
    if (s->d.world > 1 && !s->attached) return fail(s, LBM_ERR_STATE, "slab of a %d-slab lattice is not attached to its neighbours (lbm_ipc_attach)", s->d.world);
    return LBM_OK;
}

int aa_canonical(LbmSim *s) {
    const SlabParams &P = s->P;
    if (!s->scratch_dense) CU(cudaMalloc(&s->scratch_dense, sizeof(float) * 9 * (size_t)P.h * P.nx));
    dim3 block(64, 4);
    k_aa_canonical<<<grid2d(P.nx, P.h, block), block, 0, s->stream>>>(P, s->scratch_dense);
    return check_launch(s, "k_aa_canonical");
}

constexpr int kGraphSteps = 16;

// graphs that contain k_step_vec / k_step_mixed launches (their geometry follows the mask)
void invalidate_step_graphs(LbmSim *s) {
    for (auto &g : s->graph_steps)
        if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    if (s->graph_frame) { cudaGraphExecDestroy(s->graph_frame); s->graph_frame = nullptr; }
}

void invalidate_graphs(LbmSim *s) {
    invalidate_step_graphs(s);
    for (auto &g : s->graph_pairs)
        if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    for (auto &g : s->graph_pframes)
        if (g) { cudaGraphExecDestroy(g); g = nullptr; }
}

// (multi-slab launches are replayable too: the step counter the edge CTAs compare against lives on the device)
bool graphs_enabled(const LbmSim *s) { return !(s->d.flags & LBM_FLAG_NO_GRAPH); }

// tex: the macro texture to sample (nullptr = the one the latest update wrote)
int launch_particles(LbmSim *s, const __half *tex = nullptr, cudaStream_t stream = nullptr) {
    SlabParams Q = s->P;
    if (tex) Q.macro16 = const_cast<__half *>(tex);
    cudaError_t e = launch_particle_update(Q, s->field, s->pu, s->particles, s->canvas, stream ? stream : s->stream);
    if (e != cudaSuccess) return fail(s, LBM_ERR_CUDA, "launch of k_particle_update failed: %s", cudaGetErrorString(e));
    s->launches++;
    return LBM_OK;
}


// ------------------------------------------------------------------ two updates per sweep (lbm_fused.cuh)

// After a sweep the new state sits in the buffer the sweep wrote; two reference updates would have left it in
// the buffer they started from.  Exchanging the pointers restores that labelling (the swap index is unchanged).
void exchange_buffers(LbmSim *s) {
    std::swap(s->P.f[0], s->P.f[1]);
    std::swap(s->P.up[0], s->P.up[1]);
    std::swap(s->P.dn[0], s->P.dn[1]);
    std::swap(s->f_off[0], s->f_off[1]);
    s->flip ^= 1;
}

// Work items of a sweep: every strip (a warp, 60 output columns) is cut into row blocks, and each block costs two
// redundant rows of update 1.
int fuse_geometry(LbmSim *s) {
    FuseGeom &g = s->fuse;
    const int G = s->P.nx / kFuseCells;
    const int h = s->P.h;
    g.strips = (G + kFuseOut - 1) / kFuseOut;
    g.ctas_x = (g.strips + kFuseWarps - 1) / kFuseWarps;
    int fixed = 0, fixed0 = 8;
    if (const char *e = getenv("LBM_FUSE_ROWS")) fixed = atoi(e);   // uniform height (tests, tuning)
    if (const char *e = getenv("LBM_FUSE_ROWS0")) fixed0 = std::max(1, atoi(e));
    int H = fixed;
    if (H <= 0) {
        int sms = 148, per_sm = LBM_FUSE_MIN_CTAS;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
        if (s->d.world > 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_frame2<true, true, 0, false>, kFuseThreads, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_frame2<true, false, 0, false>, kFuseThreads, 0);
        // Uniform height giving at least ~6 waves of CTAs, between 8 and 32 rows (measured on B200 in rounds 1 / 2: with
        // fewer waves the sweep ended in a long tail; 4096^2 best at 16 rows, 16384^2 at 32).  Single slab, since the
        // packed last column (below) took the partially filled CTAs out of that tail: ~3 waves and up to 64 rows — taller
        // blocks, less redundant work — with the last wave cut at quarter height (4096^2 at 32 rows 132.6 -> 136.3 GLUPS,
        // 16384^2 at 64 rows +1.8 %; profiles/r02_s3_experiments.md).  Slabs keep the rule they were measured with.
        const bool tall = s->d.world == 1 && LBM_FUSE_PACK;
        // (64-row blocks paid only on 64 KB rows: 16384^2 +1.8 %, but 8192^2 porous 76.1 -> 74.6 GLUPS)
        int h_min = 8, h_max = (tall && (long long)s->P.pitch * 4 >= 65536) ? 64 : 32;
        if (const char *e = getenv("LBM_FUSE_HMIN")) h_min = std::max(1, atoi(e));
        if (const char *e = getenv("LBM_FUSE_HMAX")) h_max = std::max(h_min, atoi(e));
        int waves = tall ? 3 : 6;
        if (const char *e = getenv("LBM_FUSE_WAVES")) waves = std::max(1, atoi(e));
        const long long resident = (long long)sms * std::max(per_sm, 1);
        H = (int)((long long)h * g.ctas_x / (waves * resident));
        H = std::max(h_min, std::min(H, h_max));
        // A lattice too small to fill the GPU even once at h_min rows per block (the reference's own 600 x 375) is
        // bound by the latency of a block's serial march, (H + 2) row iterations: the shortest blocks that still fit
        // one wave (>= 2 rows: a neighbour slab reads two) finish soonest; their redundant update-1 rows cost nothing
        // while SMs idle.
        if ((long long)g.ctas_x * ((h + h_min - 1) / h_min) < resident && !getenv("LBM_FUSE_HMIN"))
            H = (int)std::max<long long>(2, ((long long)h * g.ctas_x + resident - 1) / resident);
        // A warp streams 18 planes; with blocks that do not straddle 2 MB pages it needs one page per plane.  Measured
        // on slabs of 16384 x 2048 (64 KB rows, 32 rows per page): 32-row blocks 135 GLUPS per GPU, 31-row blocks 118.
        // Snap to a power-of-two fraction of the rows per page when the row pitch allows.
        const long long row_bytes = (long long)s->P.pitch * 4, page = 2ll << 20;
        if (page % row_bytes == 0 && (s->P.plane * 4) % page == 0) {
            const int rpp = (int)(page / row_bytes);
            int best = 0;
            int top = rpp; // power-of-two fractions of a page, and whole pages when a page holds fewer rows than h_max
            while (top * 2 <= h_max) top *= 2;
            for (int c = top; c >= h_min; c >>= 1) {
                if (c > h_max || (c <= rpp ? (rpp % c) != 0 : (c % rpp) != 0)) continue;
                if (!best || std::abs(c - H) < std::abs(best - H)) best = c;
            }
            if (best) H = best;
        }
    }
    int H0 = std::min(fixed0, H); // the first strip column (the inlet column of a channel) in short blocks: see k_frame2
    if (s->d.world > 1) { H = std::max(H, 2); H0 = std::max(H0, 2); } // a neighbour reads two rows: one block must hold both
    // blocks of height `hh` in dispatch order (lbm_sweep_blocks, host_logic.cpp)
    // Tail shaping: when the grid is only a few waves long, the rows dispatched last — about one wave of CTAs — are cut
    // at half height, so that the sweep ends with short work items instead of a ragged last wave (4096^2: 6.6 waves,
    // SMs active 91 % of the launch without it).  LBM_FUSE_TAIL=0 turns it off.
    int tail_rows = 0, tail_div = (s->d.world == 1 && LBM_FUSE_PACK) ? 4 : 2;
    double tail_waves = 1.0;
    {
        int sms = 148, per_sm = LBM_FUSE_MIN_CTAS;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
        const long long resident = (long long)sms * per_sm;
        const long long blocks = (long long)g.ctas_x * ((h + H - 1) / H);
        const bool on = getenv("LBM_FUSE_TAIL") ? atoi(getenv("LBM_FUSE_TAIL")) != 0 : true;
        if (const char *e = getenv("LBM_FUSE_TAIL_WAVES")) tail_waves = atof(e);
        if (const char *e = getenv("LBM_FUSE_TAIL_DIV")) tail_div = std::max(2, atoi(e));
        if (on && fixed <= 0 && g.ctas_x > 1 && H >= 8 && blocks > resident && blocks < 24 * resident) {
            // short blocks for about tail_waves waves of CTAs
            const long long tail_blocks = (long long)(tail_waves * (double)resident / (double)(g.ctas_x - 1)) + 1;
            tail_rows = (int)(((tail_blocks * (H / tail_div) + H - 1) / H) * H);
            if (tail_rows >= h / 2) tail_rows = 0;
        }
    }
    auto cut = [&](int hh, std::vector<int2> &out) {
        const int th = std::max(hh / tail_div, 2);
        std::vector<int32_t> flat(2 * ((size_t)h / std::max(std::min(hh, th), 1) + 4));
        int32_t n_edge = 0;
        const int n = lbm_sweep_blocks_tail(h, hh, hh == H ? tail_rows : 0, th, flat.data(), (int32_t)(flat.size() / 2), &n_edge);
        for (int k = 0; k < n; k++) out.push_back(make_int2(flat[2 * k], flat[2 * k + 1]));
        return (int)n_edge;
    };
    std::vector<int2> items0, items;
    g.edge0 = cut(H0, items0);
    g.edge = cut(H, items);
    g.rowblocks0 = (int)items0.size();
    g.rowblocks = (int)items.size();
    if (g.ctas_x == 1) { // a single strip column: everything is "column 0"
        items0 = items;
        g.rowblocks0 = g.rowblocks;
        g.edge0 = g.edge;
        g.rowblocks = 0;
        g.edge = 0;
        items.clear();
    }
    items0.insert(items0.end(), items.begin(), items.end());
    if (s->d_fuse_rows) { cudaFree(s->d_fuse_rows); s->d_fuse_rows = nullptr; }
    CU(cudaMalloc(&s->d_fuse_rows, sizeof(int2) * items0.size()));
    CU(cudaMemcpy(s->d_fuse_rows, items0.data(), sizeof(int2) * items0.size(), cudaMemcpyHostToDevice));
    g.items = s->d_fuse_rows;
    g.negzero2 = 0x8000000080000000ull;
    g.pack_a = g.pack_n = 0;
#if LBM_FUSE_PACK
    // The last CTA column holds strips % kFuseWarps strips.  With one or two (4096 columns: 69 strips = 17 CTAs + ONE
    // strip) its CTAs would keep the registers of four warps for the work of one: they take four (two) row blocks instead
    // (4096^2: 126.8 -> 132.6 GLUPS, 8192^2 porous 72.8 -> 75.1, 16384^2 +1.4 %; profiles/r02_s3_experiments.md).  Not on a
    // lattice that fits the GPU in one wave (nothing competes for the slots).  On slabs too (a 16384 x 2048 slab: 140.0 ->
    // 143.4 GLUPS): the first packed CTA holds the edge row blocks, it waits for the neighbours as a whole and each of its
    // warps signals for its own row block (frame2_is_edge).
    {
        const int a = g.strips % kFuseWarps;
        const int knob = getenv("LBM_FUSE_PACK") ? atoi(getenv("LBM_FUSE_PACK")) : 1; // 0 off, 2 = also on one-wave lattices (tests)
        const bool on = knob != 0;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
        const long long resident = (long long)sms * LBM_FUSE_MIN_CTAS;
        const long long blocks = (long long)g.rowblocks0 + (long long)g.rowblocks * (g.ctas_x - 1);
        if (on && g.ctas_x >= 3 && a != 0 && kFuseWarps % a == 0 && a < kFuseWarps && (blocks > resident || knob == 2)) {
            g.pack_a = a;
            g.pack_n = (g.rowblocks + kFuseWarps / a - 1) / (kFuseWarps / a);
        }
    }
#endif
    return LBM_OK;
}

int run_ring_check(LbmSim *s) {
    CU(cudaMemsetAsync(s->d_fuse_flags + 1, 0, sizeof(unsigned int), s->stream));
    k_ring_check<<<64, 256, 0, s->stream>>>(s->P, s->d_fuse_flags + 1);
    int rc = check_launch(s, "k_ring_check");
    if (rc) return rc;
    unsigned int stale = 0;
    CU(cudaMemcpyAsync(&stale, s->d_fuse_flags + 1, sizeof(stale), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->fuse_blocked = stale != 0;
    s->ring_check_needed = false;
    return LBM_OK;
}

// The sweep finds the values the per-direction clamp changes by an unsigned compare of bit patterns against max_i
// (collide2): that needs every max_i to be >= +0 and not NaN.  Anything else takes the single-update kernels.
bool clamp_limits_ordinary(const Coef &k) {
    for (int i = 0; i < 9; i++) {
        uint32_t b;
        memcpy(&b, &k.mx[i], sizeof(b));
        if (b > 0x7f800000u) return false;
    }
    return true;
}

// May the next two updates run as one sweep?  (see the header comment of lbm_fused.cuh)
bool fuse_possible(const LbmSim *s) {
    return !s->aa && clamp_limits_ordinary(s->P.k) && !(s->d.flags & (LBM_FLAG_KERNEL_GENERIC | LBM_FLAG_NO_FUSE)) &&
           (s->d.world == 1 || s->attached) && (s->P.nx % kFuseCells) == 0 &&
           s->d.ny / s->d.world >= 4 /* the thinnest slab: every rank must answer alike */ && !s->P.macro32;
}

int fuse_eligible(LbmSim *s, bool *ok) {
    *ok = false;
    if (!fuse_possible(s) || s->countdown_left > 0) return LBM_OK;
    if (s->halo_retire_pending) {
        // Multi-slab: every armed force cell has finished its countdown by now, in the rows that own them.  The copies
        // in the info halo rows (read by the edge row blocks of a sweep) follow: material 1, like collide_stream.wgsl:58-60.
        k_retire_halo<<<(s->P.nx + 255) / 256, 256, 0, s->stream>>>(s->P, 0);
        int rc = check_launch(s, "k_retire_halo");
        if (rc) return rc;
        k_derive_halo<<<(s->P.pitch + 255) / 256, 256, 0, s->stream>>>(s->P, s->cls_halo, s->cls_halo + s->P.pitch);
        if ((rc = check_launch(s, "k_derive_halo"))) return rc;
        s->halo_retire_pending = false;
    }
    if (s->ring_check_needed) {
        if (s->d.world > 1) {
            // the verdict has to be the same on every rank, and k_ring_check only sees this slab: a write that puts a
            // solid next to the outer ring (or a restore) keeps a multi-slab lattice on single updates until lbm_reset
            s->fuse_blocked = true;
            s->ring_check_needed = false;
        } else {
            int rc = run_ring_check(s);
            if (rc) return rc;
        }
    }
    *ok = !s->fuse_blocked;
    return LBM_OK;
}

// Programmatic dependent launch (LbmSim::pdl): the sweep kernel signals at its very start that its successor in the stream
// may be launched, and waits — first thing, before it touches memory — for its predecessor to have completed
// (griddepcontrol in k_frame2).  The next sweep's CTAs are then already resident when the last CTAs of this one retire:
// the launch latency between dependent sweeps disappears (it is what a frame of a small, L2-resident lattice consists of).
template <typename K>
void launch_sweep_kernel(const LbmSim *s, K kernel, unsigned int grid, int first) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kFuseThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = s->pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, s->P, s->sync, first, s->fuse);
}

template <bool SYMW, bool SLABS, bool MASKED>
void launch_frame2m(const LbmSim *s, unsigned int grid, int first, int macro) {
    if (macro == 2) launch_sweep_kernel(s, k_frame2<SYMW, SLABS, 2, MASKED>, grid, first);
    else if (macro == 1) launch_sweep_kernel(s, k_frame2<SYMW, SLABS, 1, MASKED>, grid, first);
    else launch_sweep_kernel(s, k_frame2<SYMW, SLABS, 0, MASKED>, grid, first);
}

template <bool SYMW, bool SLABS>
void launch_frame2(const LbmSim *s, unsigned int grid, int first, int macro) {
    if (s->use_masked) launch_frame2m<SYMW, SLABS, true>(s, grid, first, macro);
    else launch_frame2m<SYMW, SLABS, false>(s, grid, first, macro);
}

// Two updates starting from buffer `first`: ONE launch, then the pointer exchange.  With a macro texture the sweep
// stores the field of t+2 into it; mid_texture: also the field of t+1 into s->macro_mid (frames with tracer particles).
int launch_pair(LbmSim *s, int first, bool mid_texture = false) {
    const FuseGeom &g = s->fuse;
    const long long blocks = (long long)g.rowblocks0 + g.pack_n + (long long)g.rowblocks * (g.ctas_x - 1 - (g.pack_n ? 1 : 0));
    if (blocks > 2147483647ll) return fail(s, LBM_ERR_INVALID_ARG, "lattice too large for one sweep launch");
    // rho * w is shared between the directions of equal weight when the uploaded weights allow it
    const Coef &k = s->P.k;
    auto same4 = [](const float *a) { return a[0] == a[1] && a[0] == a[2] && a[0] == a[3]; };
    // (also the per-direction limits: the sweep tests the largest value of each class against one limit)
    const bool symw = same4(k.w + 1) && same4(k.w + 5) && same4(k.mx + 1) && same4(k.mx + 5);
    const unsigned int grid = (unsigned int)blocks;
    int macro = s->P.macro16 ? 1 : 0;
    if (mid_texture) {
        if (!s->P.macro16 || !s->macro_mid) return fail(s, LBM_ERR_STATE, "sweep with a mid-frame texture on a handle without macro textures");
        macro = 2;
    }
    s->P.macro16_mid = s->macro_mid;
    if (s->d.world > 1) {
        if (symw) launch_frame2<true, true>(s, grid, first, macro);
        else launch_frame2<false, true>(s, grid, first, macro);
    } else {
        if (symw) launch_frame2<true, false>(s, grid, first, macro);
        else launch_frame2<false, false>(s, grid, first, macro);
    }
    int rc = check_launch(s, "k_frame2");
    if (rc) return rc;
    exchange_buffers(s);
    s->steps_since_reset += 2;
    s->macro_writes += 2;
    s->fused_sweeps++;
    s->prev_stale = true;
    return LBM_OK;
}

// The non-current buffer after a sweep holds the state two updates back.  Whoever needs what the reference has
// there (the state one update back: lbm_read_distributions(previous), the on-demand macro field, ...) gets it
// recomputed: one ordinary update from t into a temporary, copied over the stale buffer.
int materialize_prev(LbmSim *s) {
    if (!s->prev_stale) return LBM_OK;
    {
        int rc = ensure_mixed(s);
        if (rc) return rc;
    }
    if (s->d.world > 1) {
        // Multi-slab: an ordinary update t -> t+1 into the third buffer, synchronised with the neighbour slabs like any
        // other (COLLECTIVE: every rank gets here through the same read call), then that buffer takes the stale one's place.
        SlabParams &P = s->P;
        const int old = s->swap ^ 1;
        SlabParams Q = P;
        Q.f[0] = P.f[old]; Q.up[0] = P.up[old]; Q.dn[0] = P.dn[old];
        Q.f[1] = s->spare_f; Q.up[1] = s->spare_up; Q.dn[1] = s->spare_dn;
        Q.macro16 = nullptr; Q.macro32 = nullptr;
        int n = 0;
        cudaError_t e = launch_step_vec(Q, s->sync, s->mixed, 0, s->stream, &n);
        if (e != cudaSuccess) return fail(s, LBM_ERR_CUDA, "recomputing the previous buffer failed: %s", cudaGetErrorString(e));
        s->launches += n;
        std::swap(P.f[old], s->spare_f);
        std::swap(P.up[old], s->spare_up);
        std::swap(P.dn[old], s->spare_dn);
        std::swap(s->f_off[old], s->spare_off);
        invalidate_graphs(s); // buffer pointers are baked into captured launches
        s->prev_stale = false;
        return LBM_OK;
    }
    // Single slab: the same, into a spare buffer this handle keeps from its first use on (no allocation, copy or
    // host synchronisation per call), then the pointers are exchanged.
    SlabParams &P = s->P;
    const int old = s->swap ^ 1;
    const size_t bytes = sizeof(float) * 9 * P.plane;
    if (!s->prev_spare) {
        CU(cudaMalloc(&s->prev_spare_alloc, bytes));
        s->prev_spare = s->prev_spare_alloc;
        CU(cudaMemsetAsync(s->prev_spare, 0, bytes, s->stream));
    }
    // Slots no update ever writes (a solid cell's live slot whose reader sits on the outer ring, boundary.wgsl:19)
    // are constants of a buffer: the spare — a former state buffer — already holds them unless the lattice was
    // reset / restored or a solid was painted next to the ring since; only then is the stale buffer copied first.
    SlabParams Q = P;
    Q.f[0] = P.f[old]; Q.up[0] = P.up[old]; Q.dn[0] = P.dn[old];
    Q.f[1] = s->prev_spare; Q.up[1] = s->prev_spare + (size_t)(P.h - 1) * P.pitch; Q.dn[1] = s->prev_spare;
    Q.macro16 = nullptr; Q.macro32 = nullptr;
    int n = 0;
    cudaError_t e = cudaSuccess;
    if (s->spare_keep_dirty) e = cudaMemcpyAsync(s->prev_spare, P.f[old], bytes, cudaMemcpyDeviceToDevice, s->stream);
    s->spare_keep_dirty = false;
    if (e == cudaSuccess) e = launch_step_vec(Q, s->sync, s->mixed, 0, s->stream, &n);
    if (e != cudaSuccess) return fail(s, LBM_ERR_CUDA, "recomputing the previous buffer failed: %s", cudaGetErrorString(e));
    s->launches += n;
    std::swap(P.f[old], s->prev_spare);
    P.up[old] = P.f[old] + (size_t)(P.h - 1) * P.pitch;
    P.dn[old] = P.f[old];
    invalidate_graphs(s); // buffer pointers are baked into captured launches
    s->prev_stale = false;
    return LBM_OK;
}

// Captures `body` (kernel launches on s->stream) into an executable graph.
template <typename F>
int capture_graph(LbmSim *s, cudaGraphExec_t *out, uint64_t *kernels, F body) {
    const uint64_t launches = s->launches, since = s->steps_since_reset, writes = s->macro_writes, sweeps = s->fused_sweeps;
    const int64_t countdown = s->countdown_left;
    const bool stale = s->prev_stale;
    const int flip = s->flip;
    CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    int rc = body();
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(s->stream, &g);
    *kernels = s->launches - launches;
    s->launches = launches; // nothing ran yet
    s->steps_since_reset = since;
    s->macro_writes = writes;
    s->fused_sweeps = sweeps;
    s->countdown_left = countdown;
    s->prev_stale = stale;
    if (s->flip != flip) exchange_buffers(s); // a captured body must contain an even number of sweeps; be safe anyway
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(s, LBM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(s, LBM_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    return LBM_OK;
}

}  // namespace

// =================================================================== lifecycle

extern "C" int lbm_abi_version(void) { return LBM_B200_ABI_VERSION; }

extern "C" int lbm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *lbm_status_string(int st) {
    switch (st) {
        case LBM_OK: return "ok";
        case LBM_ERR_INVALID_ARG: return "invalid argument";
        case LBM_ERR_CUDA: return "CUDA error";
        case LBM_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case LBM_ERR_OUT_OF_MEMORY: return "out of device memory";
        case LBM_ERR_UNSUPPORTED: return "unsupported";
        case LBM_ERR_STATE: return "call order violated";
        default: return "unknown status";
    }
}

extern "C" const char *lbm_last_error(const LbmSim *s) { return s ? s->err.c_str() : g_create_error.c_str(); }

extern "C" void lbm_destroy(LbmSim *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    invalidate_graphs(s);
    for (int k = 0; k < 2; k++)
        if (s->peer_ipc[k] && s->peer_base[k]) cudaIpcCloseMemHandle(s->peer_base[k]);
    cudaFree(s->arena);
    cudaFree(s->P.cls);
    cudaFree(s->P.nbr);
    cudaFree(s->P.info);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    cudaFree(s->macro_buf[0] ? s->macro_buf[0] : s->P.macro16);
    cudaFree(s->macro_buf[1]);
    cudaFree(s->macro_mid);
    cudaFree(s->macro_mid2);
    cudaFree(s->macro_alt);
    if (s->side_stream) { cudaStreamSynchronize(s->side_stream); cudaStreamDestroy(s->side_stream); }
    if (s->ev_sweep) cudaEventDestroy(s->ev_sweep);
    for (auto e : s->ev_part) if (e) cudaEventDestroy(e);
    cudaFree(s->prev_spare_alloc);
    for (auto p : s->stage) if (p) cudaFreeHost(p);
    for (auto e : s->stage_ev) if (e) cudaEventDestroy(e);
    if (s->ev_ready) cudaEventDestroy(s->ev_ready);
    for (auto e : s->ev_copied) if (e) cudaEventDestroy(e);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    cudaFree(s->scratch32);
    cudaFree(s->scratch16);
    cudaFree(s->curl16);
    cudaFree(s->present32);
    cudaFree(s->scratch_dense);
    cudaFree(s->d_mass);
    cudaFree(s->d_fuse_flags);
    cudaFree(s->d_fuse_rows);
    cudaFree(s->cls_halo);
    cudaFree(s->mixed.list);
    cudaFree(s->mixed.count_dev);
    cudaFree(s->particles);
    cudaFree(s->canvas);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    cudaGetLastError(); // nothing from the teardown may surface as another handle's launch failure
    delete s;
}

static int create_impl(LbmSim *s, const LbmDesc *desc) {
    s->d = *desc;
    if (s->d.world < 1) s->d.world = 1;
    const LbmDesc &d = s->d;
    if (d.nx < 3 || d.ny < 3) return fail(s, LBM_ERR_INVALID_ARG, "lattice %dx%d too small (need >= 3x3)", d.nx, d.ny);
    if (d.rank < 0 || d.rank >= d.world) return fail(s, LBM_ERR_INVALID_ARG, "rank %d outside world %d", d.rank, d.world);
    if (d.world > 1 && d.ny / d.world < 2) return fail(s, LBM_ERR_INVALID_ARG, "ny=%d gives slabs thinner than 2 rows for world=%d", d.ny, d.world);
    if (d.lattice_pixel_size < 1) return fail(s, LBM_ERR_INVALID_ARG, "lattice_pixel_size must be >= 1");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(s, LBM_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (d.device >= ndev) return fail(s, LBM_ERR_INVALID_ARG, "device %d of %d", d.device, ndev);
    if (d.device >= 0) { CU(cudaSetDevice(d.device)); s->device = d.device; }
    else CU(cudaGetDevice(&s->device));

    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));

    SlabParams &P = s->P;
    P.nx = d.nx;
    P.ny = d.ny;
    P.y0 = (int)((int64_t)d.ny * d.rank / d.world);
    const int y1 = (int)((int64_t)d.ny * (d.rank + 1) / d.world);
    P.h = y1 - P.y0;
    P.pitch = (int)align_up((size_t)d.nx, 32);
    P.plane = align_up((size_t)P.h * P.pitch, 32);

    s->aa = (d.flags & LBM_FLAG_AA) != 0;
    if (s->aa && d.world > 1) return fail(s, LBM_ERR_UNSUPPORTED, "LBM_FLAG_AA is single-slab only");
    if (s->aa && (d.flags & LBM_FLAG_KERNEL_GENERIC)) return fail(s, LBM_ERR_UNSUPPORTED, "LBM_FLAG_AA has its own kernels");
    const size_t fbytes = align_up(sizeof(float) * 9 * P.plane, 256);
    const size_t n_buf = s->aa ? 1 : (d.world > 1 ? 3 : 2);
    s->f_off[0] = 0;
    s->f_off[1] = s->aa ? 0 : fbytes;
    s->spare_off = 2 * fbytes;
    s->flag_off = n_buf * fbytes;
    s->arena_bytes = n_buf * fbytes + 256;
    CU(cudaMalloc(&s->arena, s->arena_bytes));
    CU(cudaMemsetAsync(s->arena, 0, s->arena_bytes, s->stream));
    P.f[0] = reinterpret_cast<float *>(s->arena + s->f_off[0]);
    P.f[1] = reinterpret_cast<float *>(s->arena + s->f_off[1]);
    if (d.world > 1) s->spare_f = reinterpret_cast<float *>(s->arena + s->spare_off);
    // world == 1: rows -1 / h wrap onto the own rows h-1 / 0 (layout_and_fn.wgsl:45-49)
    for (int b = 0; b < 2; b++) {
        P.up[b] = P.f[b] + (size_t)(P.h - 1) * P.pitch;
        P.dn[b] = P.f[b];
    }
    P.up_plane = P.dn_plane = P.plane;

    const size_t cells = (size_t)P.h * P.pitch;
    CU(cudaMalloc(&P.cls, cells));
    CU(cudaMalloc(&P.nbr, cells));
    CU(cudaMemsetAsync(P.cls, CLS_SOLID, cells, s->stream));
    CU(cudaMemsetAsync(P.nbr, 0, cells, s->stream));
    CU(cudaMalloc(&s->d_fuse_flags, 2 * sizeof(unsigned int)));
    CU(cudaMemsetAsync(s->d_fuse_flags, 0, 2 * sizeof(unsigned int), s->stream));
    if (d.world > 1) {
        CU(cudaMalloc(&s->cls_halo, 2 * (size_t)P.pitch));
        CU(cudaMemsetAsync(s->cls_halo, CLS_SOLID, 2 * (size_t)P.pitch, s->stream));
        P.cls_up = s->cls_halo;
        P.cls_dn = s->cls_halo + P.pitch;
    } else { // periodic wrap: the halo rows are the own rows h-1 and 0
        P.cls_up = P.cls + (size_t)(P.h - 1) * P.pitch;
        P.cls_dn = P.cls;
    }
    const size_t info_bytes = sizeof(LatticeInfo) * (size_t)(P.h + 2) * P.nx;
    CU(cudaMalloc(&P.info, info_bytes));
    CU(cudaMemsetAsync(P.info, 0, info_bytes, s->stream));
    if (d.flags & LBM_FLAG_MACRO_EVERY_STEP) {
        CU(cudaMalloc(&P.macro16, sizeof(__half) * 4 * (size_t)P.h * P.nx));
        CU(cudaMemsetAsync(P.macro16, 0, sizeof(__half) * 4 * (size_t)P.h * P.nx, s->stream));
        s->macro_buf[0] = P.macro16;
    }
    CU(cudaMalloc(&s->d_mass, sizeof(double)));
    s->mixed.warps_per_row = (d.nx + 127) / 128;
    s->mixed.total = (uint32_t)s->mixed.warps_per_row * (uint32_t)std::max(P.h - 2, 0);
    if (s->mixed.total > 0) {
        CU(cudaMalloc(&s->mixed.list, sizeof(uint32_t) * s->mixed.total));
        CU(cudaMalloc(&s->mixed.count_dev, sizeof(uint32_t)));
    }

    s->canvas_w = d.canvas_w > 0 ? d.canvas_w : d.nx * d.lattice_pixel_size;
    s->canvas_h = d.canvas_h > 0 ? d.canvas_h : d.ny * d.lattice_pixel_size;
    if (d.max_particles > 0) {
        if (d.world > 1) return fail(s, LBM_ERR_UNSUPPORTED, "tracer particles are single-slab only");
        CU(cudaMalloc(&s->particles, sizeof(TrajectoryParticle) * (size_t)d.max_particles));
        CU(cudaMemsetAsync(s->particles, 0, sizeof(TrajectoryParticle) * (size_t)d.max_particles, s->stream));
        CU(cudaMalloc(&s->canvas, sizeof(Pixel) * (size_t)s->canvas_w * s->canvas_h));
        CU(cudaMemsetAsync(s->canvas, 0, sizeof(Pixel) * (size_t)s->canvas_w * s->canvas_h, s->stream));
    }
    // default FieldUniform (d2q9_node.rs:65-76); lbm_write_field_uniform may replace it
    lbm_field_uniform_new(d.nx, d.ny, (uint32_t)d.lattice_pixel_size, s->canvas_w, s->canvas_h, &s->field);

    if (const char *e = getenv("LBM_FUSE_MASKED")) s->use_masked = atoi(e) != 0; // A/B runs, tests
    if (const char *e = getenv("LBM_PARTICLE_OVERLAP")) s->overlap_particles = atoi(e) != 0;
    if (const char *e = getenv("LBM_PDL")) s->pdl = atoi(e) != 0;
    s->sync.flags = reinterpret_cast<unsigned int *>(s->arena + s->flag_off);
    s->sync.world = d.world;
    {
        int rc = fuse_geometry(s);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(s->stream));
    return LBM_OK;
}

extern "C" int lbm_create(const LbmDesc *desc, LbmSim **out) {
    if (!desc || !out) return fail(nullptr, LBM_ERR_INVALID_ARG, "null argument");
    if (desc->struct_size != sizeof(LbmDesc))
        return fail(nullptr, LBM_ERR_INVALID_ARG, "LbmDesc.struct_size %u != %zu", desc->struct_size, sizeof(LbmDesc));
    *out = nullptr;
    LbmSim *s = new LbmSim();
    int rc = create_impl(s, desc);
    if (rc != LBM_OK) {
        g_create_error = s->err;
        lbm_destroy(s);
        return rc;
    }
    *out = s;
    return LBM_OK;
}

// =================================================================== uploads

extern "C" int lbm_write_uniform(LbmSim *s, const LbmUniform *u) {
    if (!s || !u) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (!uniform_is_d2q9(u))
        return fail(s, LBM_ERR_UNSUPPORTED, "LbmUniform e_w_max / inversed_direction are not the D2Q9 set of fluid/mod.rs:39-52");
    if (s->prev_stale && s->have_uniform && s->have_info) {
        // the buffer one update back (and the on-demand field pulled from it) was produced with the OLD coefficients in
        // the reference: recompute it before they are replaced (collective on a multi-slab lattice)
        CU(cudaSetDevice(s->device));
        int rc = materialize_prev(s);
        if (rc) return rc;
    }
    s->u = *u;
    s->P.k.omega = u->omega;
    s->P.k.fluid_ty = u->fluid_ty;
    for (int i = 0; i < 9; i++) {
        s->P.k.w[i] = u->e_w_max[i][2];
        s->P.k.mx[i] = u->e_w_max[i][3];
    }
    s->have_uniform = true;
    invalidate_graphs(s);
    return LBM_OK;
}

extern "C" int lbm_write_field_uniform(LbmSim *s, const FieldUniform *f) {
    if (!s || !f) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (f->lattice_size[0] != s->d.nx || f->lattice_size[1] != s->d.ny)
        return fail(s, LBM_ERR_INVALID_ARG, "FieldUniform.lattice_size %dx%d does not match the handle's %dx%d",
                    f->lattice_size[0], f->lattice_size[1], s->d.nx, s->d.ny);
    s->field = *f;
    invalidate_graphs(s);
    return LBM_OK;
}

namespace {

// What the bytes of a mask write mean for the step schedule, found on the host from the caller's buffer (so that no
// device round trip is needed, and so that every rank of a multi-slab lattice — all are handed the same call —
// reaches the same verdict):
//   armed        largest block_iter > 0 of an inlet / force cell: single updates while it counts down
//   border_solid a solid within one cell of the outer ring.  Only there can a solid painted over live fluid leave a
//                value that a ring cell keeps pulling (boundary.wgsl:19 never rewrites those slots): the sweeps assume 0
//                (k_ring_check decides), and the buffer one update back must then be the reference's, not a stale one.
struct WriteScan { int32_t armed; bool border_solid; };

WriteScan scan_written_cells(int nx, int ny, uint64_t byte_offset, const void *src, uint64_t nbytes) {
    int32_t armed = 0, border = 1;
    lbm_scan_lattice_info_write(nx, ny, byte_offset, src, nbytes, &armed, &border); // host_logic.cpp (CPU-testable)
    return WriteScan{armed, border != 0};
}

// Host bytes -> device through the pinned staging ring; returns without waiting for the device.  Small writes share
// a slot (16-byte writes of a drag: one memcpy + one cudaMemcpyAsync each); a slot is reused only after the copies
// that read it have run (event recorded when the ring moves on).
int staged_upload(LbmSim *s, char *dst, const char *src, uint64_t nbytes) {
    while (nbytes > 0) {
        int k = s->stage_next;
        if (s->stage_cursor >= LbmSim::kStageBytes) { // slot full: close it, move on
            CU(cudaEventRecord(s->stage_ev[k], s->stream));
            s->stage_used[k] = true;
            k = s->stage_next = (k + 1) % LbmSim::kStageSlots;
            s->stage_cursor = 0;
            if (s->stage_used[k]) CU(cudaEventSynchronize(s->stage_ev[k]));
        }
        if (!s->stage[k]) {
            CU(cudaMallocHost(&s->stage[k], LbmSim::kStageBytes));
            CU(cudaEventCreateWithFlags(&s->stage_ev[k], cudaEventDisableTiming));
        }
        const size_t chunk = (size_t)std::min<uint64_t>(nbytes, LbmSim::kStageBytes - s->stage_cursor);
        char *stage = s->stage[k] + s->stage_cursor;
        memcpy(stage, src, chunk);
        CU(cudaMemcpyAsync(dst, stage, chunk, cudaMemcpyHostToDevice, s->stream));
        s->stage_cursor += (chunk + 15) & ~(size_t)15;
        dst += chunk; src += chunk; nbytes -= chunk;
    }
    return LBM_OK;
}

}  // namespace

extern "C" int lbm_write_lattice_info(LbmSim *s, uint64_t byte_offset, const void *src, uint64_t nbytes) {
    if (!s || (!src && nbytes)) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    const SlabParams &P = s->P;
    const uint64_t row_bytes = (uint64_t)P.nx * sizeof(LatticeInfo);
    const uint64_t total = row_bytes * (uint64_t)P.ny;
    if (byte_offset > total || nbytes > total - byte_offset)
        return fail(s, LBM_ERR_INVALID_ARG, "write of %llu bytes at %llu exceeds the %llu-byte info buffer",
                    (unsigned long long)nbytes, (unsigned long long)byte_offset, (unsigned long long)total);
    if (nbytes == 0) return LBM_OK;
    CU(cudaSetDevice(s->device));
    // Cost: O(bytes written) on the host, asynchronous on the device (d2q9_node.rs:298 issues one 16-byte write per
    // sample point of a drag; fluid_simulator.rs:158-173).
    const bool aligned = byte_offset % sizeof(LatticeInfo) == 0 && nbytes % sizeof(LatticeInfo) == 0;
    WriteScan w{0, true};
    if (aligned) w = scan_written_cells(P.nx, P.ny, byte_offset, src, nbytes);
    if (s->mask_written_since_reset && w.border_solid) {
        // a solid painted over live fluid next to the ring keeps that cell's values in BOTH buffers (ring cells may
        // pull them for ever): the buffer that is not current must hold what the reference holds there, not the
        // state of t-2 a sweep left (COLLECTIVE on a multi-slab lattice: every rank sees the same bytes)
        int rc = materialize_prev(s);
        if (rc) return rc;
        s->ring_check_needed = true;
        s->spare_keep_dirty = true;
    }
    if (w.armed > 0) {
        s->countdown_left = std::max<int64_t>(s->countdown_left, (int64_t)w.armed + 1);
        if (s->d.world > 1) s->halo_retire_pending = true;
    }
    if (!aligned && s->d.world > 1) s->fuse_blocked = true; // cannot be judged alike on every rank: single updates until lbm_reset
    const uint64_t lo = byte_offset, hi = byte_offset + nbytes;
    int touched_lo = P.h + 2, touched_hi = -1; // halo-indexed rows r = 0..h+1
    // (halo-indexed local row range, global first row) of the three pieces this slab keeps
    struct Piece { int r0, rows, gy; };
    const Piece pieces[3] = {
        {0, 1, (P.y0 - 1 + P.ny) % P.ny},
        {1, P.h, P.y0},
        {P.h + 1, 1, (P.y0 + P.h) % P.ny},
    };
    for (const Piece &pc : pieces) {
        const uint64_t g0 = (uint64_t)pc.gy * row_bytes, g1 = g0 + (uint64_t)pc.rows * row_bytes;
        const uint64_t a = std::max(lo, g0), b = std::min(hi, g1);
        if (a >= b) continue;
        char *dst = reinterpret_cast<char *>(P.info) + (uint64_t)pc.r0 * row_bytes + (a - g0);
        int rc = staged_upload(s, dst, static_cast<const char *>(src) + (a - lo), b - a);
        if (rc) return rc;
        touched_lo = std::min(touched_lo, pc.r0 + (int)((a - g0) / row_bytes));
        touched_hi = std::max(touched_hi, pc.r0 + (int)((b - 1 - g0) / row_bytes));
    }
    if (touched_hi >= 0) {
        s->have_info = true;
        // owned row l = r - 1; a changed row alters the neighbour bits of rows l-1 .. l+1
        int rc = derive_rows(s, touched_lo - 2, touched_hi + 1, !aligned && s->d.world == 1);
        if (rc) return rc;
        if (!aligned && s->d.world == 1) { // the cells cannot be parsed on the host: ask the device what is armed
            unsigned int armed = 0;
            CU(cudaMemcpyAsync(&armed, s->d_fuse_flags, sizeof(armed), cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            if (armed) s->countdown_left = std::max<int64_t>(s->countdown_left, (int64_t)armed + 1);
        }
    }
    return LBM_OK;
}

extern "C" int lbm_generate_lattice_info(LbmSim *s, int32_t kind, uint64_t seed, float solid_fraction) {
    if (!s) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (kind != FIELD_ANIMATION_POISEUILLE && kind != FIELD_ANIMATION_LID_DRIVEN_CAVITY &&
        kind != FIELD_ANIMATION_CUSTOM && kind != LBM_PRESET_POROUS)
        return fail(s, LBM_ERR_INVALID_ARG, "unknown preset %d", kind);
    CU(cudaSetDevice(s->device));
    dim3 block(64, 4);
    k_generate<<<grid2d(s->P.nx, s->P.h + 2, block), block, 0, s->stream>>>(s->P, kind, seed, solid_fraction);
    int rc = check_launch(s, "k_generate");
    if (rc) return rc;
    s->have_info = true;
    return derive_rows(s, 0, s->P.h);
}

// =================================================================== compute

extern "C" int lbm_reset(LbmSim *s) {
    if (!s) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (!s->have_uniform) return fail(s, LBM_ERR_STATE, "lbm_write_uniform has not been called");
    if (!s->have_info) return fail(s, LBM_ERR_STATE, "no lattice info uploaded or generated");
    CU(cudaSetDevice(s->device));
    dim3 block(64, 4);
    k_init<<<grid2d(s->P.nx, s->P.h, block), block, 0, s->stream>>>(s->P);
    int rc = check_launch(s, "k_init");
    if (rc) return rc;
    if (s->spare_f) CU(cudaMemsetAsync(s->spare_f, 0, sizeof(float) * 9 * s->P.plane, s->stream));
    if (s->d.world > 1) { // the neighbours' k_init retires armed force cells of rows y0-1 / y0+h (init.wgsl:51-59): so do the copies
        k_retire_halo<<<(s->P.nx + 255) / 256, 256, 0, s->stream>>>(s->P, 1);
        if ((rc = check_launch(s, "k_retire_halo"))) return rc;
    }
    s->halo_retire_pending = false;
    s->spare_keep_dirty = true;
    s->swap = 0;
    s->steps_since_reset = 0;
    s->macro_writes++; // init.wgsl:62 rewrites the texture
    // both buffers are freshly written: solids hold zeros, nothing is armed
    s->prev_stale = false;
    s->countdown_left = 0;
    s->fuse_blocked = false;
    s->ring_check_needed = false;
    s->mask_written_since_reset = false;
    // init.wgsl:51-59 may have turned armed force cells back into bulk
    rc = derive_rows(s, 0, s->P.h);
    s->mask_written_since_reset = true; // every later mask write may paint a solid over live fluid
    return rc;
}

extern "C" int lbm_step(LbmSim *s, int32_t swap_index) {
    if (!s) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (swap_index != 0 && swap_index != 1) return fail(s, LBM_ERR_INVALID_ARG, "swap_index must be 0 or 1");
    int rc = ready_to_step(s);
    if (rc) return rc;
    if (s->aa && swap_index != s->swap)
        return fail(s, LBM_ERR_STATE, "in-place (AA) state is in layout %d: the next step must use swap_index %d", s->swap, s->swap);
    CU(cudaSetDevice(s->device));
    if ((rc = ensure_mixed(s))) return rc;
    if (swap_index != s->swap && (rc = materialize_prev(s))) return rc; // stepping from the previous buffer
    rc = launch_step(s, swap_index);
    if (rc) return rc;
    s->swap = swap_index ^ 1;
    return LBM_OK;
}

// Builds (once) the graph of kGraphSteps / 2 two-update sweeps for the current pointer / swap state.
static int ensure_pair_graph(LbmSim *s) {
    const int key = s->flip * 2 + s->swap;
    if (s->graph_pairs[key]) return LBM_OK;
    const int first = s->swap;
    return capture_graph(s, &s->graph_pairs[key], &s->graph_pairs_kernels[key], [&]() {
        int r = LBM_OK;
        for (int i = 0; i < kGraphSteps / 2 && r == LBM_OK; i++) r = launch_pair(s, first);
        return r;
    });
}

static int ensure_step_graph(LbmSim *s) {
    cudaGraphExec_t &g = s->graph_steps[s->swap];
    if (g) return LBM_OK;
    const int first = s->swap;
    return capture_graph(s, &g, &s->graph_steps_kernels[s->swap], [&]() {
        int r = LBM_OK;
        for (int i = 0; i < kGraphSteps && r == LBM_OK; i++) r = launch_step(s, first ^ (i & 1));
        return r;
    });
}

extern "C" int lbm_step_n(LbmSim *s, int32_t n) {
    if (!s || n < 0) return fail(s, LBM_ERR_INVALID_ARG, "bad argument");
    int rc = ready_to_step(s);
    if (rc) return rc;
    CU(cudaSetDevice(s->device));
    int left = n;
    const bool graphs = graphs_enabled(s) && n >= 2 * kGraphSteps;
    {   // a run of sweeps needs no mixed-warp list; anything that may fall back to single updates does
        bool fz = false;
        if ((rc = fuse_eligible(s, &fz))) return rc;
        if ((!fz || (n & 1)) && (rc = ensure_mixed(s))) return rc;
    }
    // decide and capture before the timed region starts (the common case: nothing is counting down)
    bool fused = false;
    if (left >= 2 && s->countdown_left == 0) {
        if ((rc = fuse_eligible(s, &fused))) return rc;
        if (graphs && (rc = fused ? ensure_pair_graph(s) : ensure_step_graph(s))) return rc;
    }
    CU(cudaEventRecord(s->ev0, s->stream));
    if (s->countdown_left > 0 && fuse_possible(s)) {
        // force cells that are still counting down mutate the info buffer between updates: single updates
        while (left > 0 && s->countdown_left > 0) {
            if ((rc = launch_step(s, s->swap))) return rc;
            s->swap ^= 1;
            left--;
        }
        if (left >= 2 && (rc = fuse_eligible(s, &fused))) return rc;
    }
    if (fused) {
        // two updates per launch; the swap index is the same after every sweep
        if (graphs && left >= 2 * kGraphSteps) {
            if ((rc = ensure_pair_graph(s))) return rc;
            const int key = s->flip * 2 + s->swap;
            for (; left >= kGraphSteps; left -= kGraphSteps) { // an even number of sweeps: pointers unchanged
                CU(cudaGraphLaunch(s->graph_pairs[key], s->stream));
                s->launches += s->graph_pairs_kernels[key];
                s->steps_since_reset += kGraphSteps;
                s->macro_writes += kGraphSteps;
                s->fused_sweeps += kGraphSteps / 2;
                s->prev_stale = true;
            }
        }
        for (; left >= 2; left -= 2)
            if ((rc = launch_pair(s, s->swap))) return rc;
    } else if (graphs && left >= 2 * kGraphSteps) {
        if ((rc = ensure_step_graph(s))) return rc;
        cudaGraphExec_t g = s->graph_steps[s->swap];
        for (; left >= kGraphSteps; left -= kGraphSteps) { // an even number of steps: swap index unchanged
            CU(cudaGraphLaunch(g, s->stream));
            s->launches += s->graph_steps_kernels[s->swap];
            s->steps_since_reset += kGraphSteps;
            s->macro_writes += kGraphSteps;
            s->prev_stale = false;
            s->countdown_left = std::max<int64_t>(0, s->countdown_left - kGraphSteps);
        }
    }
    for (; left > 0; left--) {
        if ((rc = launch_step(s, s->swap))) return rc;
        s->swap ^= 1;
    }
    CU(cudaEventRecord(s->ev1, s->stream));
    s->timed = true;
    return LBM_OK;
}

// FluidSimulator::compute (fluid_simulator.rs:217-232) n_frames times:
// step(0), particle update, step(1), particle update.
extern "C" int lbm_compute_frames(LbmSim *s, int32_t n_frames) {
    if (!s || n_frames < 0) return fail(s, LBM_ERR_INVALID_ARG, "bad argument");
    int rc = ready_to_step(s);
    if (rc) return rc;
    if (s->swap != 0)
        return fail(s, LBM_ERR_STATE, "lbm_compute_frames starts with step(0): the current state must be in buffer 0 (an even number of updates since lbm_reset)");
    const bool with_particles = s->particles != nullptr;
    if (with_particles) {
        if (!s->have_pu) return fail(s, LBM_ERR_STATE, "lbm_write_particle_uniform has not been called");
        if (!s->P.macro16) return fail(s, LBM_ERR_STATE, "tracer particles need LBM_FLAG_MACRO_EVERY_STEP");
    }
    CU(cudaSetDevice(s->device));
    auto frame = [&]() {
        int r = launch_step(s, 0);
        if (r == LBM_OK && with_particles) r = launch_particles(s);
        if (r == LBM_OK) r = launch_step(s, 1);
        if (r == LBM_OK && with_particles) r = launch_particles(s);
        return r;
    };
    // A frame is one two-update sweep whenever nothing mutates between its updates.  Tracer particles only READ the
    // field (particle_update.wgsl:11,20): the sweep stores the texture of t+1 and of t+2, and the two particle passes
    // run after it — same inputs, same order among themselves as fluid_simulator.rs:223-231.
    bool fused = false;
    if (n_frames > 0 && (rc = fuse_eligible(s, &fused))) return rc;
    if (!fused && (rc = ensure_mixed(s))) return rc;
    if (fused && with_particles) {
        if (!s->macro_mid) {
            const size_t bytes = sizeof(__half) * 4 * (size_t)s->P.h * s->P.nx;
            CU(cudaMalloc(&s->macro_mid, bytes));
            CU(cudaMemsetAsync(s->macro_mid, 0, bytes, s->stream));
        }
        auto pframe = [&]() {
            int r = launch_pair(s, 0, true);
            if (r == LBM_OK) r = launch_particles(s, s->macro_mid);
            if (r == LBM_OK) r = launch_particles(s);
            return r;
        };
        int left = n_frames;
        const bool graphs = graphs_enabled(s) && n_frames >= 4;
        // A captured run of F frames (F even: the buffer pointers and the texture set end where they started).  Frame f
        // stores its textures into set (f + 1) & 1, so the last frame's are the handle's own; the particle passes of frame
        // f wait for sweep f, run one after the other on the side stream (the order among the passes is the reference's),
        // and sweep f + 2 — the next writer of their texture set — waits for them.
        const int F = n_frames >= 8 ? 8 : 4;
        const int gkey = s->flip * 2 + (F == 8 ? 1 : 0);
        if (graphs && !s->graph_pframes[gkey]) {
            const bool overlap = s->overlap_particles;
            if (overlap && !s->side_stream) {
                const size_t bytes = sizeof(__half) * 4 * (size_t)s->P.h * s->P.nx;
                CU(cudaMalloc(&s->macro_mid2, bytes));
                CU(cudaMalloc(&s->macro_alt, bytes));
                CU(cudaMemsetAsync(s->macro_mid2, 0, bytes, s->stream));
                CU(cudaMemsetAsync(s->macro_alt, 0, bytes, s->stream));
                CU(cudaEventCreateWithFlags(&s->ev_sweep, cudaEventDisableTiming));
                for (auto &e : s->ev_part) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                CU(cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking));
            }
            rc = capture_graph(s, &s->graph_pframes[gkey], &s->graph_pframes_kernels[gkey], [&]() {
                int r = LBM_OK;
                if (!overlap) {
                    for (int f = 0; f < F && r == LBM_OK; f++) r = pframe();
                    return r;
                }
                __half *const mid[2] = {s->macro_mid, s->macro_mid2}, *const end[2] = {s->P.macro16, s->macro_alt};
                cudaError_t e = cudaSuccess;
                for (int f = 0; f < F && r == LBM_OK && e == cudaSuccess; f++) {
                    const int set = (f + 1) & 1;
                    if (f >= 2) e = cudaStreamWaitEvent(s->stream, s->ev_part[set], 0); // passes of frame f - 2 read this set
                    s->P.macro16 = end[set];
                    s->macro_mid = mid[set];
                    r = launch_pair(s, 0, true);
                    if (e == cudaSuccess) e = cudaEventRecord(s->ev_sweep, s->stream);
                    if (e == cudaSuccess) e = cudaStreamWaitEvent(s->side_stream, s->ev_sweep, 0);
                    if (r == LBM_OK && e == cudaSuccess) r = launch_particles(s, mid[set], s->side_stream);
                    if (r == LBM_OK && e == cudaSuccess) r = launch_particles(s, end[set], s->side_stream);
                    if (e == cudaSuccess) e = cudaEventRecord(s->ev_part[set], s->side_stream);
                }
                s->P.macro16 = end[0];
                s->macro_mid = mid[0];
                // join: the passes of the last frame (set 0) are the last work of the side stream
                if (e == cudaSuccess) e = cudaStreamWaitEvent(s->stream, s->ev_part[0], 0);
                if (r == LBM_OK && e != cudaSuccess) r = fail(s, LBM_ERR_CUDA, "capturing overlapped particle frames failed: %s", cudaGetErrorString(e));
                return r;
            });
            if (rc) return rc;
        }
        CU(cudaEventRecord(s->ev0, s->stream));
        if (graphs) {
            for (; left >= F; left -= F) {
                CU(cudaGraphLaunch(s->graph_pframes[gkey], s->stream));
                s->launches += s->graph_pframes_kernels[gkey];
                s->steps_since_reset += 2 * F;
                s->macro_writes += 2 * F;
                s->fused_sweeps += F;
                s->prev_stale = true;
            }
        }
        for (; left > 0; left--)
            if ((rc = pframe())) return rc;
        CU(cudaEventRecord(s->ev1, s->stream));
        s->timed = true;
        return LBM_OK;
    }
    if (fused) { // no tracer particles: a frame is one launch
        int left = n_frames;
        const bool graphs = graphs_enabled(s) && n_frames >= kGraphSteps;
        if (graphs && (rc = ensure_pair_graph(s))) return rc;
        CU(cudaEventRecord(s->ev0, s->stream));
        if (graphs) {
            const int key = s->flip * 2 + s->swap;
            for (; left >= kGraphSteps / 2; left -= kGraphSteps / 2) {
                CU(cudaGraphLaunch(s->graph_pairs[key], s->stream));
                s->launches += s->graph_pairs_kernels[key];
                s->steps_since_reset += kGraphSteps;
                s->macro_writes += kGraphSteps;
                s->fused_sweeps += kGraphSteps / 2;
                s->prev_stale = true;
            }
        }
        for (; left > 0; left--)
            if ((rc = launch_pair(s, 0))) return rc;
        CU(cudaEventRecord(s->ev1, s->stream));
        s->timed = true;
        return LBM_OK;
    }
    const bool use_graph = graphs_enabled(s) && n_frames >= 4;
    if (use_graph && !s->graph_frame) {
        rc = capture_graph(s, &s->graph_frame, &s->graph_frame_kernels, frame);
        if (rc) return rc;
    }
    CU(cudaEventRecord(s->ev0, s->stream));
    for (int f = 0; f < n_frames; f++) {
        if (use_graph) {
            CU(cudaGraphLaunch(s->graph_frame, s->stream));
            s->launches += s->graph_frame_kernels;
            s->steps_since_reset += 2;
            s->macro_writes += 2;
            s->prev_stale = false;
            s->countdown_left = std::max<int64_t>(0, s->countdown_left - 2);
        } else {
            rc = frame();
            if (rc) return rc;
        }
    }
    if (n_frames > 0) s->swap = 0; // a frame ends with step(1): the next step reads buffer 0
    CU(cudaEventRecord(s->ev1, s->stream));
    s->timed = true;
    return LBM_OK;
}

extern "C" int lbm_swap_index(const LbmSim *s) { return s ? s->swap : -1; }

extern "C" int lbm_refresh_previous(LbmSim *s) {
    if (!s) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    CU(cudaSetDevice(s->device));
    return materialize_prev(s);
}

extern "C" int lbm_sync(LbmSim *s) {
    if (!s) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    if (s->copy_stream) {
        CU(cudaStreamSynchronize(s->copy_stream));
        s->copy_pending[0] = s->copy_pending[1] = false;
    }
    if (s->d.world > 1) {
        // sticky word set by an edge CTA whose wait for a neighbour slab ran into the 4 s bound
        unsigned int timed_out = 0;
        CU(cudaMemcpy(&timed_out, s->sync.flags + 3, sizeof(timed_out), cudaMemcpyDeviceToHost));
        if (timed_out)
            return fail(s, LBM_ERR_STATE, "a step of slab %d gave up waiting for a neighbour slab (is every slab being stepped?); results are invalid", s->d.rank);
    }
    return LBM_OK;
}

// =================================================================== read-back / restore

extern "C" int lbm_slab_rows(const LbmSim *s, int32_t *y0, int32_t *rows) {
    if (!s) return LBM_ERR_INVALID_ARG;
    if (y0) *y0 = s->P.y0;
    if (rows) *rows = s->P.h;
    return LBM_OK;
}

extern "C" int lbm_read_distributions(LbmSim *s, int32_t which, float *dst) {
    if (!s || !dst || (which != 0 && which != 1)) return fail(s, LBM_ERR_INVALID_ARG, "bad argument");
    CU(cudaSetDevice(s->device));
    if (which != s->swap) { // the previous buffer: recompute it if the last launch was a two-update sweep
        int rc = materialize_prev(s);
        if (rc) return rc;
    }
    const SlabParams &P = s->P;
    const size_t n = (size_t)P.h * P.nx;
    if (s->aa) {
        if (which != s->swap) return fail(s, LBM_ERR_UNSUPPORTED, "in-place (AA) handle: only the current state (which = %d) exists", s->swap);
        if (s->swap == 1) { // shifted layout: canonicalise to the reference layout first
            int rc = aa_canonical(s);
            if (rc) return rc;
            CU(cudaMemcpyAsync(dst, s->scratch_dense, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            return LBM_OK;
        }
    }
    for (int i = 0; i < 9; i++)
        CU(cudaMemcpy2DAsync(dst + i * n, sizeof(float) * P.nx, P.f[which] + i * P.plane, sizeof(float) * P.pitch,
                             sizeof(float) * P.nx, P.h, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return LBM_OK;
}

extern "C" int lbm_write_distributions(LbmSim *s, int32_t which, const float *src) {
    if (!s || !src || (which != 0 && which != 1)) return fail(s, LBM_ERR_INVALID_ARG, "bad argument");
    if (s->aa && (which != 0 || s->swap != 0))
        return fail(s, LBM_ERR_UNSUPPORTED, "in-place (AA) handle: distributions can only be written as buffer 0 in the natural layout (after lbm_reset or an even number of steps)");
    CU(cudaSetDevice(s->device));
    {   // keep the other buffer what the reference would hold there before this one is replaced
        int rc = materialize_prev(s);
        if (rc) return rc;
    }
    const SlabParams &P = s->P;
    const size_t n = (size_t)P.h * P.nx;
    for (int i = 0; i < 9; i++)
        CU(cudaMemcpy2DAsync(P.f[which] + i * P.plane, sizeof(float) * P.pitch, src + i * n, sizeof(float) * P.nx,
                             sizeof(float) * P.nx, P.h, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->ring_check_needed = true; // restored solids may hold values a ring cell pulls (k_ring_check)
    s->spare_keep_dirty = true;
    return LBM_OK;
}

extern "C" int lbm_read_macro(LbmSim *s, int32_t format, void *dst) {
    if (!s || !dst) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (format != LBM_MACRO_F32_PLANES && format != LBM_MACRO_RGBA16F) return fail(s, LBM_ERR_INVALID_ARG, "unknown macro format %d", format);
    CU(cudaSetDevice(s->device));
    const SlabParams &P = s->P;
    const size_t n = (size_t)P.h * P.nx;
    if (format == LBM_MACRO_RGBA16F && P.macro16) {
        // written by the step itself (LBM_FLAG_MACRO_EVERY_STEP); right after lbm_read_macro_async switched
        // textures the newest field is still in the other one
        const __half *newest = (s->macro_writes == s->macro_writes_at_flip) ? s->macro_buf[s->macro_cur ^ 1] : P.macro16;
        CU(cudaMemcpyAsync(dst, newest, sizeof(__half) * 4 * n, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        return LBM_OK;
    }
    if (s->aa && s->steps_since_reset != 0)
        return fail(s, LBM_ERR_UNSUPPORTED, "in-place (AA) handle: the step overwrites its inputs, create it with LBM_FLAG_MACRO_EVERY_STEP to read the field");
    // On demand: the buffer the last step read is still intact (A/B ping-pong), so pulling from it
    // again yields exactly the (rho, u) that step computed (collide_stream.wgsl:43-74).
    int rc = ready_to_step(s);
    if (rc) return rc;
    if ((rc = materialize_prev(s))) return rc; // after a two-update sweep that buffer is recomputed first
    SlabParams Q = P;
    if (format == LBM_MACRO_F32_PLANES) {
        rc = ensure_scratch32(s);
        if (rc) return rc;
        Q.macro32 = s->scratch32;
        Q.macro16 = nullptr;
    } else {
        rc = ensure_scratch16(s);
        if (rc) return rc;
        Q.macro16 = s->scratch16;
        Q.macro32 = nullptr;
    }
    if (s->steps_since_reset == 0) {
        k_macro_after_init<<<148 * 8, 256, 0, s->stream>>>(Q);
        rc = check_launch(s, "k_macro_after_init");
    } else {
        dim3 block(64, 4);
        k_step_generic<1><<<grid2d(P.nx, P.h, block), block, 0, s->stream>>>(Q, s->swap ^ 1, 0, P.h);
        rc = check_launch(s, "k_step_generic<macro>");
    }
    if (rc) return rc;
    if (format == LBM_MACRO_F32_PLANES)
        CU(cudaMemcpyAsync(dst, Q.macro32, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, s->stream));
    else
        CU(cudaMemcpyAsync(dst, Q.macro16, sizeof(__half) * 4 * n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return LBM_OK;
}

namespace {

// curl_update.wgsl over the newest macro texture (the step's own texture, or the on-demand one) into s->curl16.
// *tex_out: the macro texture it was computed from.
int compute_curl(LbmSim *s, const __half **tex_out) {
    const SlabParams &P = s->P;
    const size_t n = (size_t)P.h * P.nx;
    const __half *tex = nullptr;
    if (P.macro16) {
        tex = (s->macro_writes == s->macro_writes_at_flip) ? s->macro_buf[s->macro_cur ^ 1] : P.macro16;
    } else {
        // no per-step texture: produce it like lbm_read_macro(LBM_MACRO_RGBA16F) does
        if (s->aa && s->steps_since_reset != 0)
            return fail(s, LBM_ERR_UNSUPPORTED, "in-place (AA) handle: create it with LBM_FLAG_MACRO_EVERY_STEP to read derived fields");
        int rc = ready_to_step(s);
        if (rc) return rc;
        if ((rc = materialize_prev(s))) return rc;
        if ((rc = ensure_scratch16(s))) return rc;
        SlabParams Q = P;
        Q.macro16 = s->scratch16;
        Q.macro32 = nullptr;
        if (s->steps_since_reset == 0) {
            k_macro_after_init<<<148 * 8, 256, 0, s->stream>>>(Q);
            rc = check_launch(s, "k_macro_after_init");
        } else {
            dim3 block(64, 4);
            k_step_generic<1><<<grid2d(P.nx, P.h, block), block, 0, s->stream>>>(Q, s->swap ^ 1, 0, P.h);
            rc = check_launch(s, "k_step_generic<macro>");
        }
        if (rc) return rc;
        tex = s->scratch16;
    }
    if (!s->curl16) CU(cudaMalloc(&s->curl16, sizeof(__half) * 4 * n));
    dim3 block(64, 4);
    k_curl<<<grid2d(P.nx, P.h, block), block, 0, s->stream>>>(tex, P.nx, P.h, s->curl16);
    int rc = check_launch(s, "k_curl");
    if (rc) return rc;
    *tex_out = tex;
    return LBM_OK;
}

}  // namespace

extern "C" int lbm_read_curl(LbmSim *s, void *dst) {
    if (!s || !dst) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (s->d.world > 1) return fail(s, LBM_ERR_UNSUPPORTED, "lbm_read_curl is single-slab only");
    CU(cudaSetDevice(s->device));
    const __half *tex = nullptr;
    int rc = compute_curl(s, &tex);
    if (rc) return rc;
    CU(cudaMemcpyAsync(dst, s->curl16, sizeof(__half) * 4 * (size_t)s->P.h * s->P.nx, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return LBM_OK;
}

// lbm/present.wgsl over the newest macro texture and its curl: fragment outputs of canvas rows [row0, row0 + rows).
extern "C" int lbm_read_present(LbmSim *s, int32_t row0, int32_t rows, float *dst) {
    if (!s || !dst) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (s->d.world > 1) return fail(s, LBM_ERR_UNSUPPORTED, "lbm_read_present is single-slab only");
    const int W = s->field.canvas_size[0], H = s->field.canvas_size[1];
    if (W <= 0 || H <= 0) return fail(s, LBM_ERR_STATE, "FieldUniform.canvas_size is %dx%d", W, H);
    if (row0 < 0 || rows < 0 || row0 > H || rows > H - row0)
        return fail(s, LBM_ERR_INVALID_ARG, "rows [%d, %d + %d) outside the %d rows of the canvas", row0, row0, rows, H);
    if (rows == 0) return LBM_OK;
    CU(cudaSetDevice(s->device));
    const __half *tex = nullptr;
    int rc = compute_curl(s, &tex);
    if (rc) return rc;
    const size_t n = (size_t)rows * W;
    if (s->present_cap < n) {
        CU(cudaStreamSynchronize(s->stream));
        cudaFree(s->present32);
        s->present32 = nullptr;
        s->present_cap = 0;
        CU(cudaMalloc(&s->present32, sizeof(float4) * n));
        s->present_cap = n;
    }
    dim3 block(64, 4);
    k_present<<<grid2d(W, rows, block), block, 0, s->stream>>>(tex, s->curl16, s->P.nx, s->P.h, W, H, row0, rows, s->present32);
    if ((rc = check_launch(s, "k_present"))) return rc;
    CU(cudaMemcpyAsync(dst, s->present32, sizeof(float4) * n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return LBM_OK;
}

// Pipelined read-back of the RGBA16F macro texture: the copy runs on its own stream while the handle
// goes on stepping into a second texture.  dst must stay valid (and should be pinned) until lbm_sync.
extern "C" int lbm_read_macro_async(LbmSim *s, void *dst) {
    if (!s || !dst) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (!s->P.macro16) return fail(s, LBM_ERR_STATE, "lbm_read_macro_async needs LBM_FLAG_MACRO_EVERY_STEP");
    CU(cudaSetDevice(s->device));
    const size_t bytes = sizeof(__half) * 4 * (size_t)s->P.h * s->P.nx;
    if (!s->copy_stream) {
        CU(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s->ev_ready, cudaEventDisableTiming));
        for (auto &e : s->ev_copied) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaMalloc(&s->macro_buf[1], bytes));
    }
    const int cur = s->macro_cur, nxt = cur ^ 1;
    CU(cudaEventRecord(s->ev_ready, s->stream));              // texture `cur` is final once the queued steps ran
    CU(cudaStreamWaitEvent(s->copy_stream, s->ev_ready, 0));
    CU(cudaMemcpyAsync(dst, s->macro_buf[cur], bytes, cudaMemcpyDeviceToHost, s->copy_stream));
    CU(cudaEventRecord(s->ev_copied[cur], s->copy_stream));
    s->copy_pending[cur] = true;
    // later steps write the other texture, once the copy that may still read it has finished
    if (s->copy_pending[nxt]) CU(cudaStreamWaitEvent(s->stream, s->ev_copied[nxt], 0));
    s->macro_cur = nxt;
    s->P.macro16 = s->macro_buf[nxt];
    s->macro_writes_at_flip = s->macro_writes;
    invalidate_graphs(s); // the texture pointer is baked into captured launches
    return LBM_OK;
}

extern "C" int lbm_read_lattice_info(LbmSim *s, LatticeInfo *dst) {
    if (!s || !dst) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    CU(cudaSetDevice(s->device));
    const SlabParams &P = s->P;
    CU(cudaMemcpyAsync(dst, P.info + P.nx, sizeof(LatticeInfo) * (size_t)P.h * P.nx, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return LBM_OK;
}

extern "C" int lbm_total_mass(LbmSim *s, int32_t which, double *out) {
    if (!s || !out || (which != 0 && which != 1)) return fail(s, LBM_ERR_INVALID_ARG, "bad argument");
    CU(cudaSetDevice(s->device));
    CU(cudaMemsetAsync(s->d_mass, 0, sizeof(double), s->stream));
    int rc;
    if (which != s->swap && (rc = materialize_prev(s))) return rc;
    if (s->aa && which != s->swap) return fail(s, LBM_ERR_UNSUPPORTED, "in-place (AA) handle: only the current state (which = %d) exists", s->swap);
    if (s->aa && s->swap == 1) {
        if ((rc = aa_canonical(s))) return rc;
        k_sum_dense<<<148 * 4, 256, 0, s->stream>>>(s->scratch_dense, (size_t)9 * s->P.h * s->P.nx, s->d_mass);
        rc = check_launch(s, "k_sum_dense");
    } else {
        k_mass<<<148 * 8, 256, 0, s->stream>>>(s->P, which, s->d_mass);
        rc = check_launch(s, "k_mass");
    }
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, s->d_mass, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return LBM_OK;
}

// =================================================================== tracer particles

extern "C" int lbm_write_particle_uniform(LbmSim *s, const ParticleUniform *pu) {
    if (!s || !pu) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (pu->num[0] < 0 || pu->num[1] < 0 || (int64_t)pu->num[0] * pu->num[1] > (int64_t)s->d.max_particles)
        return fail(s, LBM_ERR_INVALID_ARG, "particle grid %dx%d exceeds max_particles=%d", pu->num[0], pu->num[1], s->d.max_particles);
    s->pu = *pu;
    s->have_pu = true;
    invalidate_graphs(s);
    return LBM_OK;
}

extern "C" int lbm_particles_write(LbmSim *s, const TrajectoryParticle *src, uint64_t count) {
    if (!s || !src) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (count > (uint64_t)s->d.max_particles) return fail(s, LBM_ERR_INVALID_ARG, "%llu particles exceed max_particles=%d", (unsigned long long)count, s->d.max_particles);
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(s->particles, src, sizeof(TrajectoryParticle) * count, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->n_particles = count;
    return LBM_OK;
}

extern "C" int lbm_particles_update(LbmSim *s) {
    if (!s) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (!s->particles) return fail(s, LBM_ERR_STATE, "handle was created with max_particles = 0");
    if (!s->have_pu) return fail(s, LBM_ERR_STATE, "lbm_write_particle_uniform has not been called");
    if (!s->P.macro16) return fail(s, LBM_ERR_STATE, "tracer particles read the macro texture: create the handle with LBM_FLAG_MACRO_EVERY_STEP");
    CU(cudaSetDevice(s->device));
    return launch_particles(s);
}

extern "C" int lbm_particles_read(LbmSim *s, TrajectoryParticle *dst, uint64_t count) {
    if (!s || !dst) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (count > (uint64_t)s->d.max_particles) return fail(s, LBM_ERR_INVALID_ARG, "count exceeds max_particles");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(dst, s->particles, sizeof(TrajectoryParticle) * count, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return LBM_OK;
}

extern "C" int lbm_canvas_clear(LbmSim *s) {
    if (!s) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (!s->canvas) return fail(s, LBM_ERR_STATE, "handle was created with max_particles = 0");
    CU(cudaSetDevice(s->device));
    CU(cudaMemsetAsync(s->canvas, 0, sizeof(Pixel) * (size_t)s->canvas_w * s->canvas_h, s->stream));
    return LBM_OK;
}

extern "C" int lbm_canvas_fade(LbmSim *s) {
    if (!s) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (!s->canvas) return fail(s, LBM_ERR_STATE, "handle was created with max_particles = 0");
    if (!s->have_pu) return fail(s, LBM_ERR_STATE, "lbm_write_particle_uniform has not been called");
    CU(cudaSetDevice(s->device));
    k_canvas_fade<<<148 * 8, 256, 0, s->stream>>>(s->canvas, (size_t)s->canvas_w * s->canvas_h, s->pu.fade_out_factor);
    return check_launch(s, "k_canvas_fade");
}

extern "C" int lbm_canvas_read(LbmSim *s, Pixel *dst) {
    if (!s || !dst) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (!s->canvas) return fail(s, LBM_ERR_STATE, "handle was created with max_particles = 0");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(dst, s->canvas, sizeof(Pixel) * (size_t)s->canvas_w * s->canvas_h, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return LBM_OK;
}

// =================================================================== multi-GPU wiring

extern "C" int lbm_ipc_export(LbmSim *s, LbmIpcBlob *out) {
    if (!s || !out) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    CU(cudaSetDevice(s->device));
    memset(out, 0, sizeof(*out));
    IpcPayload p{};
    p.magic = kIpcMagic;
    p.pid = (int32_t)getpid_portable();
    p.device = s->device;
    p.rank = s->d.rank;
    p.world = s->d.world;
    p.nx = s->P.nx; p.ny = s->P.ny; p.y0 = s->P.y0; p.h = s->P.h; p.pitch = s->P.pitch;
    p.plane = s->P.plane;
    p.f_off[0] = s->f_off[0]; p.f_off[1] = s->f_off[1];
    p.spare_off = s->spare_off;
    p.flag_off = s->flag_off;
    p.arena_bytes = s->arena_bytes;
    p.local_ptr = (uint64_t)(uintptr_t)s->arena;
    CU(cudaIpcGetMemHandle(&p.handle, s->arena));
    memcpy(out->bytes, &p, sizeof(p));
    return LBM_OK;
}

extern "C" int lbm_ipc_attach(LbmSim *s, const LbmIpcBlob *up, const LbmIpcBlob *down) {
    if (!s || !up || !down) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (s->d.world < 2) return fail(s, LBM_ERR_STATE, "a single-slab lattice has no neighbours to attach");
    if (s->attached) return fail(s, LBM_ERR_STATE, "already attached");
    CU(cudaSetDevice(s->device));
    const LbmIpcBlob *blobs[2] = {up, down};
    const int want_rank[2] = {(s->d.rank - 1 + s->d.world) % s->d.world, (s->d.rank + 1) % s->d.world};
    char *base[2] = {nullptr, nullptr};
    IpcPayload pl[2];
    for (int k = 0; k < 2; k++) {
        memcpy(&pl[k], blobs[k]->bytes, sizeof(IpcPayload));
        const IpcPayload &p = pl[k];
        if (p.magic != kIpcMagic) return fail(s, LBM_ERR_INVALID_ARG, "neighbour blob %d is not an LbmIpcBlob", k);
        if (p.rank != want_rank[k] || p.world != s->d.world || p.nx != s->P.nx || p.ny != s->P.ny || p.pitch != s->P.pitch)
            return fail(s, LBM_ERR_INVALID_ARG, "neighbour blob %d describes rank %d of %d (%dx%d), expected rank %d of %d (%dx%d)",
                        k, p.rank, p.world, p.nx, p.ny, want_rank[k], s->d.world, s->P.nx, s->P.ny);
        if (p.pid == (int32_t)getpid_portable()) {
            // same process: plain peer access
            if (p.device != s->device) {
                int can = 0;
                CU(cudaDeviceCanAccessPeer(&can, s->device, p.device));
                if (!can) return fail(s, LBM_ERR_UNSUPPORTED, "device %d cannot access peer device %d", s->device, p.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(p.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(s, LBM_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", p.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            base[k] = reinterpret_cast<char *>((uintptr_t)p.local_ptr);
        } else if (k == 1 && pl[0].pid == p.pid && pl[0].rank == p.rank && s->peer_ipc[0]) {
            base[k] = static_cast<char *>(s->peer_base[0]); // world == 2: both neighbours are the same slab
        } else {
            void *ptr = nullptr;
            CU(cudaIpcOpenMemHandle(&ptr, p.handle, cudaIpcMemLazyEnablePeerAccess));
            s->peer_base[k] = ptr;
            s->peer_ipc[k] = true;
            base[k] = static_cast<char *>(ptr);
        }
    }
    SlabParams &P = s->P;
    for (int b = 0; b < 2; b++) {
        // up neighbour's last row; down neighbour's first row
        P.up[b] = reinterpret_cast<float *>(base[0] + pl[0].f_off[b]) + (size_t)(pl[0].h - 1) * pl[0].pitch;
        P.dn[b] = reinterpret_cast<float *>(base[1] + pl[1].f_off[b]);
    }
    s->spare_up = reinterpret_cast<float *>(base[0] + pl[0].spare_off) + (size_t)(pl[0].h - 1) * pl[0].pitch;
    s->spare_dn = reinterpret_cast<float *>(base[1] + pl[1].spare_off);
    P.up_plane = pl[0].plane;
    P.dn_plane = pl[1].plane;
    s->sync.peer_flags[0] = reinterpret_cast<unsigned int *>(base[0] + pl[0].flag_off);
    s->sync.peer_flags[1] = reinterpret_cast<unsigned int *>(base[1] + pl[1].flag_off);
    s->attached = true;
    return LBM_OK;
}

// =================================================================== introspection

extern "C" uint64_t lbm_launch_count(const LbmSim *s) { return s ? s->launches : 0; }
extern "C" uint64_t lbm_fused_sweep_count(const LbmSim *s) { return s ? s->fused_sweeps : 0; }
extern "C" int lbm_sweep_uses_masked_path(const LbmSim *s) { return s && s->use_masked ? 1 : 0; }

extern "C" int lbm_edge_wait_stats(LbmSim *s, uint64_t *total_ns, uint64_t *n_waits) {
    if (!s || !total_ns || !n_waits) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    CU(cudaSetDevice(s->device));
    unsigned long long v[2] = {0, 0};
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemcpy(&v[0], s->sync.flags + 6, sizeof(v[0]), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&v[1], s->sync.flags + 8, sizeof(v[1]), cudaMemcpyDeviceToHost));
    *total_ns = v[0];
    *n_waits = v[1];
    return LBM_OK;
}

extern "C" int lbm_last_step_n_ms(LbmSim *s, float *ms) {
    if (!s || !ms) return fail(s, LBM_ERR_INVALID_ARG, "null argument");
    if (!s->timed) return fail(s, LBM_ERR_STATE, "lbm_step_n has not been called");
    CU(cudaSetDevice(s->device));
    CU(cudaEventSynchronize(s->ev1));
    CU(cudaEventElapsedTime(ms, s->ev0, s->ev1));
    return LBM_OK;
}

extern "C" void *lbm_stream(LbmSim *s) { return s ? (void *)s->stream : nullptr; }
