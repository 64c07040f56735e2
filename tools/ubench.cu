// Micro-benchmarks of the sm_100a instruction mixes the K2 sweep kernels are built from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void __launch_bounds__(512) bench(float *out, unsigned *gbuf, long long *cyc, int zero) {
    __shared__ unsigned sm[512 * 12];
    float f[8]; unsigned u[8]; unsigned long long p[8]; double dd[8];
    const double da = 1.0 + zero;
    const int tid = threadIdx.x, lane = tid & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) dd[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = tid * 0.001f + i; u[i] = tid * 7 + i + zero; p[i] = ((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] + 1.f); }
    for (int i = tid; i < 512 * 12; i += 512) sm[i] = i;
    const float a = 1.0001f + zero, b = 0.5f + zero;
    const unsigned long long pa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    const unsigned long long pb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {        // FFMA x8
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __fmaf_rn(f[i], a, b);
        } else if (MODE == 1) { // FFMA2 x8
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], pa, pb);
        } else if (MODE == 2) { // IADD3 x8
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
        } else if (MODE == 3) { // FFMA x8 + IADD x8
#pragma unroll
            for (int i = 0; i < 8; ++i) { f[i] = __fmaf_rn(f[i], a, b); asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7])); }
        } else if (MODE == 4) { // FFMA2 x8 + IADD x8
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = ffma2(p[i], pa, pb); asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7])); }
        } else if (MODE == 5) { // FFMA2 x8 + IADD x16
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = ffma2(p[i], pa, pb); asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[(i + 3) & 7]) : "r"(u[(i + 5) & 7])); }
        } else if (MODE == 6) { // SHFL x8
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = __shfl_up_sync(0xffffffffu, u[i], 1);
        } else if (MODE == 7) { // LDS.32 x8 (conflict-free)
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = sm[(u[i] & 7) * 512 + tid];
        } else if (MODE == 8) { // LDS.128 x8
#pragma unroll
            for (int i = 0; i < 8; i += 4) { uint4 v = *reinterpret_cast<uint4 *>(&sm[((u[i] & 1) * 512 + tid) * 4]); u[i] += v.x; u[i + 1] += v.y; u[i + 2] += v.z; u[i + 3] += v.w; }
        } else if (MODE == 9) { // REDUX.OR x8
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = __reduce_or_sync(0xffffffffu, u[i]) + lane;
        } else if (MODE == 10) { // VOTE.ballot x8
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] += __ballot_sync(0xffffffffu, (u[i] & 1) != 0);
        } else if (MODE == 11) { // ATOMS.ADD x8 conflict-free, no return
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(&sm[(i + (it & 3)) * 512 + tid], 1u);
        } else if (MODE == 12) { // STS.32 x8
#pragma unroll
            for (int i = 0; i < 8; ++i) sm[((it + i) & 7) * 512 + tid] = u[i];
        } else if (MODE == 13) { // global RED, 4 of 32 lanes active, L2-resident distinct addresses
#pragma unroll
            for (int i = 0; i < 8; ++i) if ((lane & 7) == 0) atomicAdd(&gbuf[((blockIdx.x * 512 + tid) * 8 + i) * 8 % (1 << 22)], 1u);
        } else if (MODE == 14) { // FFMA x12 + 10 ALU (approx. sweep mix)
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __fmaf_rn(f[i], a, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) f[i] = __fmaf_rn(f[i], a, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[0]) : "r"(u[5])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[1]) : "r"(u[6]));
        } else if (MODE == 15) { // FFMA2 x6 + FADD + 10 ALU
#pragma unroll
            for (int i = 0; i < 6; ++i) p[i] = ffma2(p[i], pa, pb);
            f[0] += f[1];
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[0]) : "r"(u[5])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[1]) : "r"(u[6]));
        } else if (MODE == 16) { // LDS.32 x4 + SHFL x4
#pragma unroll
            for (int i = 0; i < 4; ++i) { u[i] = sm[(u[i] & 7) * 512 + tid]; u[i + 4] = __shfl_up_sync(0xffffffffu, u[i + 4], 1); }
        } else if (MODE == 18) { // F2F.F64.F32 x8 (result folded back through the integer bits)
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double d = (double)f[i]; f[i] = __int_as_float(__double2hiint(d) ^ __double2loint(d)); }
        } else if (MODE == 19) { // DADD x8
#pragma unroll
            for (int i = 0; i < 8; ++i) dd[i] = __dadd_rn(dd[i], da);
        } else if (MODE == 20) { // FMUL + F2F + DADD x8: the exact-item inner step (essentia dotProduct arithmetic)
#pragma unroll
            for (int i = 0; i < 8; ++i) dd[i] = __dadd_rn(dd[i], (double)__fmul_rn(f[i], a));
        } else if (MODE == 21) { // FMUL + integer widening (3 ALU) + DADD x8
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned bts = __float_as_uint(__fmul_rn(f[i], a));
                dd[i] = __dadd_rn(dd[i], __hiloint2double((int)((bts >> 3) + 0x38000000u), (int)(bts << 29)));
            }
        } else if (MODE == 17) { // byte store to global: 32 contiguous bytes per warp
#pragma unroll
            for (int i = 0; i < 8; ++i) reinterpret_cast<unsigned char *>(gbuf)[(size_t)(blockIdx.x * 16 + (tid >> 5)) * 65536 + ((it * 8 + i) & 2047) * 32 + lane] = (unsigned char)u[i];
        }
    }
    long long t1 = clock64();
    float s = 0; unsigned us = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += f[i] + (float)dd[i]; us += u[i] + (unsigned)p[i] + (unsigned)(p[i] >> 32); }
    out[blockIdx.x * 512 + tid] = s + us + sm[tid];
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, int per_iter, float *out, unsigned *gbuf, long long *cyc) {
    bench<MODE><<<148, 512>>>(out, gbuf, cyc, 0);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<148, 512>>>(out, gbuf, cyc, 0);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double winstr = 16.0 * ITERS * per_iter;          // warp instructions per SM
    printf("%-34s %8.3f ms  %10.0f cyc  %6.3f warp-instr/clk/SM  (%5.3f /clk/SMSP) err=%s\n", name, ms, c, winstr / c, winstr / c / 4,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *out; unsigned *gbuf; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&gbuf, (size_t)256 << 20); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(gbuf, 0, (size_t)256 << 20);
    run<0>("FFMA x8", 8, out, gbuf, cyc);
    run<1>("FFMA2 x8", 8, out, gbuf, cyc);
    run<2>("IADD x8", 8, out, gbuf, cyc);
    run<3>("FFMA x8 + IADD x8", 16, out, gbuf, cyc);
    run<4>("FFMA2 x8 + IADD x8", 16, out, gbuf, cyc);
    run<5>("FFMA2 x8 + IADD/XOR x16", 24, out, gbuf, cyc);
    run<6>("SHFL x8", 8, out, gbuf, cyc);
    run<7>("LDS.32 x8", 8, out, gbuf, cyc);
    run<8>("LDS.128 x2", 2, out, gbuf, cyc);
    run<9>("REDUX.OR x8 (+IADD)", 16, out, gbuf, cyc);
    run<10>("VOTE.ballot x8 (+misc)", 8, out, gbuf, cyc);
    run<11>("ATOMS.ADD x8", 8, out, gbuf, cyc);
    run<12>("STS.32 x8", 8, out, gbuf, cyc);
    run<13>("RED.global 4/32 lanes x8", 8, out, gbuf, cyc);
    run<14>("FFMA x12 + ALU x10", 22, out, gbuf, cyc);
    run<15>("FFMA2 x6 + FADD + ALU x10", 17, out, gbuf, cyc);
    run<16>("LDS.32 x4 + SHFL x4", 8, out, gbuf, cyc);
    run<17>("STG.U8 32B/warp x8", 8, out, gbuf, cyc);
    run<18>("F2F.F64.F32 x8 (+2 ALU)", 8, out, gbuf, cyc);
    run<19>("DADD x8", 8, out, gbuf, cyc);
    run<20>("FMUL + F2F + DADD x8", 8, out, gbuf, cyc);
    run<21>("FMUL + 3 ALU widen + DADD x8", 8, out, gbuf, cyc);
    return 0;
}
