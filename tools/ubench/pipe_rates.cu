// pipe_rates.cu — issue rates of the f32 instruction forms the sweep kernel is built from, per SM sub-partition
// (B200, sm_100a).  Each kernel runs ILP independent dependency chains per thread for ITERS iterations with 8 warps
// per sub-partition; reports warp-instructions per cycle per sub-partition.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
#define ILP 8
#define ITERS 4096

template <int OP>
__global__ void __launch_bounds__(1024) k(float *out, float a0, float b0, unsigned long long nz, long long *cyc) {
    float x[ILP], y[ILP];
    f2 p[ILP], q[ILP];
    const float b = b0, c = a0;
    f2 pb, pc = nz;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b0), "f"(b0));
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        x[i] = a0 + i + threadIdx.x;
        y[i] = a0 * i;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[i]), "f"(y[i]));
        q[i] = p[i] ^ 0x0000100000001000ull;
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            if (OP == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(b), "f"(c));
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
            if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            if (OP == 6) asm volatile("mul.rn.f32 %0, %0, 0f3F800001;" : "+f"(x[i]));
            if (OP == 7) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y[i]));           // 2 distinct regs
            if (OP == 8) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));         // 2 distinct pairs
            if (OP == 9) { // mix: FADD (fma pipe) + FMNMX (alu pipe)
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c));
            }
            if (OP == 10) { // mix: FADD2 + FMNMX
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c));
            }
            if (OP == 11) { // mix: FADD2 + FMUL
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c));
            }
            if (OP == 12) { // mix: FADD2 + 2 FMNMX
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            }
            if (OP == 13) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            if (OP == 14) { unsigned int u = __float_as_uint(x[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(__float_as_uint(b)), "r"(__float_as_uint(c))); x[i] = __uint_as_float(u); }
            if (OP == 15) { unsigned int u = __float_as_uint(x[i]); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u) : "r"(__float_as_uint(b)), "r"(__float_as_uint(c))); x[i] = __uint_as_float(u); }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + y[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)q[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter) {
    float *out; long long *cyc;
    const int blocks = 148, threads = 1024; // 32 warps per SM = 8 per sub-partition
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<OP><<<blocks, threads>>>(out, 1.0f, 1.0000001f, 0x8000000080000000ull, cyc);
    k<OP><<<blocks, threads>>>(out, 1.0f, 1.0000001f, 0x8000000080000000ull, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += h[i];
    avg /= blocks;
    const double warp_instr_per_smsp = 8.0 * ITERS * ILP * per_iter;
    printf("%-28s %8.0f cycles  %.3f warp-instr/cycle/SMSP  (%.2f cycles per instr)\n", name, avg, warp_instr_per_smsp / avg, avg / warp_instr_per_smsp);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FADD r,r(shared)", 1);
    run<7>("FADD r,r(distinct)", 1);
    run<1>("FMUL r,r", 1);
    run<6>("FMUL r,imm", 1);
    run<2>("FFMA r,r,r", 1);
    run<3>("FADD2 (shared b)", 1);
    run<8>("FADD2 (distinct)", 1);
    run<13>("FMUL2", 1);
    run<4>("FFMA2 r,r,r", 1);
    run<5>("FMNMX", 1);
    run<14>("LOP3", 1);
    run<15>("IMAD", 1);
    run<9>("FADD + FMNMX", 2);
    run<10>("FADD2 + FMNMX", 2);
    run<11>("FADD2 + FMUL", 2);
    run<12>("FADD2 + 2 FMNMX", 3);
    return 0;
}
