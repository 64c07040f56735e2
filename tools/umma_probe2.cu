// Second probe for the K2 tensor sweeps: (1) A operand in TMEM (tcgen05.st + .kind::i8 TS form) vs shared memory,
// (2) cycles per MMA for N = 32 / 64 / 128, aligned vs tap-shifted descriptors.  Pair-tile scheme: every MMA is
// [plane_t | plane_t+1] x [plane'_t | plane'_t+1], 30 MMAs per block.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe2 tools/umma_probe2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int M = 128, AF = 144, BFMAX = 272;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)8 << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_i8(int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

struct Smem {
    alignas(128) uint8_t a[3][AF][16];
    alignas(128) uint8_t b[3][BFMAX][16];
    alignas(128) uint8_t zero[256][16];
    unsigned long long bar;
    uint32_t tmem_base;
};
// MODE 0: SS, tap-shifted descriptors (the real thing); 1: SS, every descriptor at tap 0 (aligned; wrong result, timing only);
// 2: TS (A tiles in TMEM columns 384 + 8 * tile)
template <int N, int MODE>
__global__ void __launch_bounds__(128) probe(const uint8_t *__restrict__ ga, const uint8_t *__restrict__ gb, int *__restrict__ out, int reps,
                                             long long *__restrict__ cycles) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem &s = *reinterpret_cast<Smem *>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 3 * AF * 16; i += 128) (&s.a[0][0][0])[i] = ga[i];
    for (int i = tid; i < 3 * BFMAX * 16; i += 128) (&s.b[0][0][0])[i] = gb[i];
    for (int i = tid; i < 256 * 16; i += 128) (&s.zero[0][0])[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s.bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    constexpr uint32_t ACOL = 384;                                     // A tiles: 15 x 8 columns
    if (MODE == 2) {
        // thread = row m: tile (pl, tp) = [frame m + 2 tp | frame m + 2 tp + 1] of plane pl
        for (int pl = 0; pl < 3; ++pl)
            for (int tp = 0; tp < 5; ++tp) {
                uint32_t v[8];
                const uint32_t *f0 = reinterpret_cast<const uint32_t *>(&s.a[pl][tid + 2 * tp][0]);
                for (int i = 0; i < 8; ++i) v[i] = f0[i];             // 32 contiguous bytes = two frames
                tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + ACOL + 8 * (pl * 5 + tp), v);
            }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    constexpr uint32_t PA = AF * 16, PB = BFMAX * 16;
    const uint32_t a0 = smem_u32(&s.a[0][0][0]), b0 = smem_u32(&s.b[0][0][0]), Z = smem_u32(&s.zero[0][0]);
    // one MMA: planes (pa, pb), tap pair tp, accumulator acc
    auto one = [&](uint32_t pa, uint32_t pb, uint32_t acc, int tp, uint32_t accum) {
        const uint32_t sa = a0 + pa * PA, sb = b0 + pb * PB;
        const int t = MODE == 1 ? 0 : 2 * tp;
        const uint64_t bd = make_desc(sb + 16 * t, tp == 4 ? Z - (sb + 16 * t) : 16);
        if (MODE == 2) mma_ts(tmem + acc * N, tmem + ACOL + 8 * (pa * 5 + tp), bd, idesc_i8(N), accum);
        else mma_ss(tmem + acc * N, make_desc(sa + 16 * t, 16), bd, idesc_i8(N), accum);
    };
    // interleaved issue: consecutive MMAs go to different accumulators (acc2 has 15, acc1 10, acc0 5)
    auto block = [&]() {
#pragma unroll
        for (int tp = 0; tp < 5; ++tp) {
            one(1, 1, 2, tp, tp != 0);
            one(0, 1, 1, tp, tp != 0);
            one(0, 2, 2, tp, 1);
            one(0, 0, 0, tp, tp != 0);
            one(2, 0, 2, tp, 1);
            one(1, 0, 1, tp, 1);
        }
    };
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (tid == 0) {
            block();
            mma_commit(&s.bar);
        }
        mbar_wait(&s.bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (out) {
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < N; c += 16) {
                int v[16];
                tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + a * N + c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int i = 0; i < 16; ++i) out[(a * M + tid) * N + c + i] = v[i];
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

std::vector<uint8_t> ha(3 * AF * 16), hb(3 * BFMAX * 16);
uint8_t *da, *db; int *dout; long long *dc;

template <int N, int MODE>
void run(const char *name, bool check) {
    cudaFuncSetAttribute(probe<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (check) {
        cudaMemset(dout, 0xff, 3 * M * N * 4);
        probe<N, MODE><<<1, 128, sizeof(Smem)>>>(da, db, dout, 1, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
        std::vector<int> got(3 * M * N);
        cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
        auto A = [&](int pl, int f, int k) { return (int)ha[(pl * AF + f) * 16 + k]; };
        auto B = [&](int pl, int f, int k) { return (int)hb[(pl * BFMAX + f) * 16 + k]; };
        long bad[3] = {0, 0, 0};
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                long long s[3] = {0, 0, 0};
                for (int t = 0; t < 9; ++t)
                    for (int k = 0; k < 16; ++k) {
                        s[0] += A(0, m + t, k) * B(0, n + t, k);
                        s[1] += A(0, m + t, k) * B(1, n + t, k) + A(1, m + t, k) * B(0, n + t, k);
                        s[2] += A(1, m + t, k) * B(1, n + t, k) + A(0, m + t, k) * B(2, n + t, k) + A(2, m + t, k) * B(0, n + t, k);
                    }
                for (int a = 0; a < 3; ++a) bad[a] += got[(a * M + m) * N + n] != (int)s[a];
            }
        printf("%s check: mismatches %ld %ld %ld of %d\n", name, bad[0], bad[1], bad[2], M * N);
    }
    for (int grid : {1, 148}) {
        const int reps = 200;
        probe<N, MODE><<<grid, 128, sizeof(Smem)>>>(da, db, nullptr, reps, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s timing: %s\n", name, cudaGetErrorString(e)); exit(1); }
        std::vector<long long> cyc(grid);
        cudaMemcpy(cyc.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
        printf("%s grid %3d: %.0f cycles per block of 30 MMAs = %.1f per MMA, %.2f cells/clk/SM\n", name, grid, (double)mx / reps,
               (double)mx / reps / 30, (double)M * N * reps / mx);
    }
}

int main() {
    srand(7);
    for (auto &x : ha) x = rand() & 255;
    for (auto &x : hb) x = rand() & 255;
    cudaMalloc(&da, ha.size()); cudaMalloc(&db, hb.size()); cudaMalloc(&dout, 3 * M * 256 * 4); cudaMalloc(&dc, 1024 * 8);
    cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice);
    run<64, 0>("SS N=64 shifted", true);
    run<80, 0>("SS N=80 shifted", true);
    run<96, 0>("SS N=96 shifted", true);
    run<112, 0>("SS N=112 shifted", true);
    run<128, 0>("SS N=128 shifted", true);
    run<160, 0>("SS N=160 shifted", true);
    run<64, 2>("TS N=64", true);
    run<80, 2>("TS N=80", true);
    return 0;
}
