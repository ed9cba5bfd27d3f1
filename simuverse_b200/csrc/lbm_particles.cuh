// lbm_particles.cuh — tracer-particle advection through the macroscopic field.
// Restates particle_update.wgsl:55-88, its src_3f/update_canvas helpers (:15-53) and
// func/bilinear_interpolate_3f.wgsl:1-12.  One thread per particle; the field is the
// RGBA16F macro texture the step writes (f16-quantised, like textureLoad in the reference).
#pragma once

#include <stdlib.h>

#include "lbm_device.cuh"

namespace lbm {

struct Tex3 { float x, y, z; };

__device__ __forceinline__ Tex3 src_3f(const __half *macro16, int nx, int ny, int u, int v) {
    const int nu = min(max(u, 0), nx - 1); // particle_update.wgsl:16-17
    const int nv = min(max(v, 0), ny - 1);
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(macro16) + ((size_t)nv * nx + nu));
    const __half2 a = *reinterpret_cast<const __half2 *>(&t.x);
    const __half2 b = *reinterpret_cast<const __half2 *>(&t.y);
    Tex3 r;
    r.x = __low2float(a);
    r.y = __high2float(a);
    r.z = __low2float(b);
    return r;
}

// WGSL f32 -> i32: truncate toward zero, saturating (cvt.rzi.s32.f32 saturates; NaN -> 0)
__device__ __forceinline__ int f2i(float v) { return __float2int_rz(v); }

__global__ void __launch_bounds__(256) k_particle_update(const __half *__restrict__ macro16, int nx, int ny,
                                                         const __grid_constant__ FieldUniform field,
                                                         const __grid_constant__ ParticleUniform pu, int poiseuille,
                                                         TrajectoryParticle *particles, Pixel *canvas) {
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int gy = blockIdx.y * blockDim.y + threadIdx.y;
    if (gx >= pu.num[0] || gy >= pu.num[1]) return;
    const size_t p_index = (size_t)gx + (size_t)gy * pu.num[0];
    // 24-byte record, 8-byte aligned: three 64-bit accesses each way
    TrajectoryParticle p;
    {
        const float2 *src = reinterpret_cast<const float2 *>(particles + p_index);
        const float2 a = src[0], b = src[1], c = src[2];
        p.pos[0] = a.x; p.pos[1] = a.y; p.pos_initial[0] = b.x; p.pos_initial[1] = b.y; p.life_time = c.x; p.fade = c.y;
    }
    if (p.life_time <= 0.1f) {
        p.fade = 0.0f;
        p.pos[0] = p.pos_initial[0];
        p.pos[1] = p.pos_initial[1];
        p.life_time = pu.life_time;
    } else {
        p.life_time = fsub(p.life_time, 1.0f);
        if (p.fade < 1.0f) {
            if (p.fade < 0.95f) p.fade = fadd(p.fade, 0.1f); else p.fade = 1.0f;
        }
        const float ijx = fsub(fdiv(p.pos[0], field.lattice_pixel_size[0]), 0.5f);
        const float ijy = fsub(fdiv(p.pos[1], field.lattice_pixel_size[1]), 0.5f);
        const int minX = f2i(floorf(ijx)), minY = f2i(floorf(ijy));
        const float fx = fsub(ijx, (float)minX), fy = fsub(ijy, (float)minY);
        const Tex3 a = src_3f(macro16, nx, ny, minX, minY);
        const Tex3 b = src_3f(macro16, nx, ny, minX, minY + 1);
        const Tex3 c = src_3f(macro16, nx, ny, minX + 1, minY);
        const Tex3 d = src_3f(macro16, nx, ny, minX + 1, minY + 1);
        const float wa = fmul(fsub(1.0f, fx), fsub(1.0f, fy));
        const float wb = fmul(fsub(1.0f, fx), fy);
        const float wc = fmul(fx, fsub(1.0f, fy));
        const float wd = fmul(fx, fy);
        const float vx = fadd(fadd(fadd(fmul(a.x, wa), fmul(b.x, wb)), fmul(c.x, wc)), fmul(d.x, wd));
        const float vy = fadd(fadd(fadd(fmul(a.y, wa), fmul(b.y, wb)), fmul(c.y, wc)), fmul(d.y, wd));
        p.pos[0] = fadd(p.pos[0], fmul(vx, pu.speed_factor));
        p.pos[1] = fadd(p.pos[1], fmul(vy, pu.speed_factor));
        // update_canvas (particle_update.wgsl:32-53); concurrent writers to one pixel race in the
        // reference too (last writer wins)
        const float speed = fadd(fabsf(vx), fabsf(vy));
        const bool skip = (!poiseuille && speed < 0.0f) || (poiseuille && speed < 0.015f);
        if (!skip && canvas) {
            const int px = f2i(p.pos[0]) - pu.point_size / 2;
            const int py = f2i(p.pos[1]) - pu.point_size / 2;
            Pixel pix;
            pix.alpha = p.fade; pix.velocity_x = vx; pix.velocity_y = vy;
            for (int dx = 0; dx < pu.point_size; dx++)
                for (int dy = 0; dy < pu.point_size; dy++) {
                    const int cx = px + dx, cy = py + dy;
                    if (cx >= 0 && cx < field.canvas_size[0] && cy >= 0 && cy < field.canvas_size[1])
                        canvas[(size_t)cx + (size_t)field.canvas_size[0] * cy] = pix;
                }
        }
    }
    {
        float2 *dst = reinterpret_cast<float2 *>(particles + p_index);
        dst[0] = make_float2(p.pos[0], p.pos[1]);
        // pos_initial never changes
        dst[2] = make_float2(p.life_time, p.fade);
    }
}

// curl_update.wgsl:12-33 — derived field of the macro texture (the reference's `_curl_cal_node`).  One thread per
// texel; taps left / top are clamped to 0, right / bottom to lattice_size, i.e. ONE PAST the last texel (:21,24),
// where textureLoad is out of bounds and wgpu returns zeros (naga ReadZeroSkipWrite).  f16 texels convert to f32
// exactly; the three additions and `curl * 3.5 + 0.5` are single IEEE operations in source order.
__global__ void __launch_bounds__(256) k_curl(const __half *__restrict__ macro16, int nx, int ny, __half *__restrict__ curl16) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx || y >= ny) return;
    auto tap = [&](int u, int v, int comp) -> float {
        if (u < 0 || u >= nx || v < 0 || v >= ny) return 0.0f;
        return __half2float(macro16[4 * ((size_t)v * nx + u) + comp]);
    };
    const int rx = min(x + 1, nx), lx = max(x - 1, 0), ty = max(y - 1, 0), by = min(y + 1, ny);
    float curl = fsub(tap(rx, y, 1), tap(lx, y, 1));
    curl = fadd(curl, tap(x, ty, 0));
    curl = fsub(curl, tap(x, by, 0));
    const __half2 a = __floats2half2_rn(fadd(fmul(curl, 3.5f), 0.5f), 0.0f);
    uint2 t;
    t.x = *reinterpret_cast<const uint32_t *>(&a);
    t.y = 0u;
    reinterpret_cast<uint2 *>(curl16)[(size_t)y * nx + x] = t;
}

// lbm/present.wgsl:21-46 — colour present of the field (fragment shader of the reference's `render_node`,
// fluid_simulator.rs:69-87; built there, its draw call commented out at :243-244).  One thread per pixel of the
// canvas_size target: position = pixel centre, uv = position / target size; macro and curl textures sampled through
// `bilinear_sampler` (ClampToEdge, linear; util/load_texture.rs:229-243) with WebGPU's filter formula in f32, one
// rounding per operation, taps summed in the order 00, 10, 01, 11 (hardware samplers use fixed-point weights of
// unspecified width, so no two implementations are bit-comparable; this arithmetic is what include/lbm_b200.h defines
// and the tests pin on the executed shader text).
// colour = hsv2rgb(curl.x, 0.6 + speed * 1.4, 0.6 + rho * 0.33) (func/color_space_convert.wgsl:2-9), alpha = rho.
// Bound by the 16 B per pixel it writes (one coalesced 128-bit store per thread); the 2 x 4 taps come out of L1 / L2.
struct BilinearTaps { int o00, o10, o01, o11; float w00, w10, w01, w11; };

__device__ __forceinline__ BilinearTaps bilinear_taps(int nx, int ny, float u, float v) {
    const float cx = fsub(fmul(u, (float)nx), 0.5f), cy = fsub(fmul(v, (float)ny), 0.5f);
    const float fx0 = floorf(cx), fy0 = floorf(cy);
    const float fx = fsub(cx, fx0), fy = fsub(cy, fy0);
    const int ix = (int)fx0, iy = (int)fy0;
    const int x0 = min(max(ix, 0), nx - 1), x1 = min(max(ix + 1, 0), nx - 1);
    const int y0 = min(max(iy, 0), ny - 1), y1 = min(max(iy + 1, 0), ny - 1);
    const float gx = fsub(1.0f, fx), gy = fsub(1.0f, fy);
    BilinearTaps t;
    t.o00 = y0 * nx + x0; t.o10 = y0 * nx + x1; t.o01 = y1 * nx + x0; t.o11 = y1 * nx + x1;
    t.w00 = fmul(gx, gy); t.w10 = fmul(fx, gy); t.w01 = fmul(gx, fy); t.w11 = fmul(fx, fy);
    return t;
}

__device__ __forceinline__ float bilinear_mix(const BilinearTaps &t, float a, float b, float c, float d) {
    float s = fmul(t.w00, a);
    s = fadd(s, fmul(t.w10, b));
    s = fadd(s, fmul(t.w01, c));
    return fadd(s, fmul(t.w11, d));
}

__device__ __forceinline__ float hsv2rgb_channel(float h, float K, float s, float v) {
    const float t = fadd(h, K);
    const float p = fabsf(fsub(fmul(fsub(t, floorf(t)), 6.0f), 3.0f)); // |fract(h + K) * 6 - 3|
    const float c = fminf(fmaxf(fsub(p, 1.0f), 0.0f), 1.0f);
    return fmul(v, fadd(fsub(1.0f, s), fmul(c, s)));                  // v * mix(1, c, s); 1 * (1 - s) is exact
}

__global__ void __launch_bounds__(256) k_present(const __half *__restrict__ macro16, const __half *__restrict__ curl16, int nx,
                                                 int ny, int W, int H, int row0, int rows, float4 *__restrict__ out) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= W || r >= rows) return;
    const int py = row0 + r;
    const float u = fdiv(fadd((float)px, 0.5f), (float)W), v = fdiv(fadd((float)py, 0.5f), (float)H);
    const BilinearTaps t = bilinear_taps(nx, ny, u, v);
    // macro texel = (u.x, u.y, rho, 1): three channels are used; of the curl texel only x
    const uint2 *m = reinterpret_cast<const uint2 *>(macro16);
    auto unpack = [](uint2 q, float &a, float &b, float &c) {
        const __half2 lo = *reinterpret_cast<const __half2 *>(&q.x), hi = *reinterpret_cast<const __half2 *>(&q.y);
        a = __low2float(lo); b = __high2float(lo); c = __low2float(hi);
    };
    float ax, ay, az, bx, by, bz, cx, cy, cz, dx, dy, dz;
    unpack(m[t.o00], ax, ay, az); unpack(m[t.o10], bx, by, bz); unpack(m[t.o01], cx, cy, cz); unpack(m[t.o11], dx, dy, dz);
    const float ux = bilinear_mix(t, ax, bx, cx, dx), uy = bilinear_mix(t, ay, by, cy, dy), rho = bilinear_mix(t, az, bz, cz, dz);
    const float curl = bilinear_mix(t, __half2float(curl16[4 * (size_t)t.o00]), __half2float(curl16[4 * (size_t)t.o10]),
                                    __half2float(curl16[4 * (size_t)t.o01]), __half2float(curl16[4 * (size_t)t.o11]));
    const float speed = fadd(fabsf(ux), fabsf(uy));
    const float sat = fadd(0.6f, fmul(speed, 1.4f)), val = fadd(0.6f, fmul(rho, 0.33f));
    float4 o;
    o.x = hsv2rgb_channel(curl, 1.0f, sat, val);
    o.y = hsv2rgb_channel(curl, fdiv(2.0f, 3.0f), sat, val);
    o.z = hsv2rgb_channel(curl, fdiv(1.0f, 3.0f), sat, val);
    o.w = rho;
    out[(size_t)r * W + px] = o;
}

// present.wgsl:19-22,43-49 — the in-place fade of the canvas that the present pass performs
__global__ void __launch_bounds__(256) k_canvas_fade(Pixel *canvas, size_t n, float fade_out_factor) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float a = canvas[i].alpha;
        if (a > 0.001f) canvas[i].alpha = (a >= 0.2f) ? fmul(a, fade_out_factor) : fmul(a, 0.5f);
    }
}

inline cudaError_t launch_particle_update(const SlabParams &P, const FieldUniform &field, const ParticleUniform &pu,
                                          TrajectoryParticle *particles, Pixel *canvas, cudaStream_t stream) {
    if (pu.num[0] <= 0 || pu.num[1] <= 0) return cudaSuccess;
    // (16, 16) threads per block like the reference's workgroups (particle_update.wgsl:55).  Other warp tiles of the particle
    // grid (4 x 8, 8 x 4, 32 x 1; LBM_PARTICLE_BLOCK_X) were measured on 1000 x 1000 tracers over the 8192^2 porous field:
    // 51.7 - 53.8 us per pass whatever the shape — the pass is bound by the latency of its scattered texel fetches and
    // canvas splats, 6 % of a frame (profiles/r02_masked_path_ab.md).
    static int bx = 0;
    if (!bx) {
        const char *e = getenv("LBM_PARTICLE_BLOCK_X");
        bx = e ? atoi(e) : 16;
        if (bx != 4 && bx != 8 && bx != 16 && bx != 32) bx = 16;
    }
    dim3 block(bx, 256 / bx);
    dim3 grid((pu.num[0] + block.x - 1) / block.x, (pu.num[1] + block.y - 1) / block.y);
    k_particle_update<<<grid, block, 0, stream>>>(P.macro16, P.nx, P.h, field, pu, P.k.fluid_ty == 0, particles, canvas);
    return cudaGetLastError();
}

}  // namespace lbm
