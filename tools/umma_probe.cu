// Probe of the tcgen05 (kind::i8) building block of the K2 tensor sweeps, standalone (no library needed):
//   - operands are byte planes of 16-byte padded frames; a row of the "stacked" operand (9 consecutive frames) is never
//     materialised: the shared-memory descriptor walks the plane with LBO = 16 B (next tap) and SBO = 128 B
//     (next 8 rows), i.e. core matrix (rows r0..r0+7, tap t) = the 128 contiguous bytes at frame r0 + t;
//   - three int32 accumulators in TMEM (limb products of equal weight), read back with tcgen05.ld 32x32b;
//   - checks every value against a CPU loop, then times MMA issue and TMEM reads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int TAPS = 9, M = 128, N = 64;
constexpr int AF = 144, BF = 80;                 // frames per plane held in shared memory (>= M + 9, N + 9)
constexpr int NPL = 3;                           // limb planes: h, l1, l2

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: ((8, m), 2) : ((16 B, SBO), LBO)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
    return d;
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W1:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D1;\n"
        "bra W1;\n"
        "D1:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

// instruction descriptor: S32 accumulate, U8 x U8, K-major both, M = 128, N
__host__ __device__ constexpr uint32_t idesc_i8(int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

struct Smem {
    alignas(128) uint8_t a[NPL][AF][16];         // owned side planes: h, l1, l2
    alignas(128) uint8_t b[NPL][BF][16];         // streamed side planes, stored l2, l1, h (so every needed LBO is positive)
    alignas(128) uint8_t zero[N][16];
    unsigned long long bar;
    uint32_t tmem_base;
};

// all MMAs of one 128 x N block: acc0 = hh, acc1 = h.l1 + l1.h, acc2 = l1.l1 + h.l2 + l2.h
__device__ __forceinline__ void issue_block(Smem *s, uint32_t tmem, uint32_t idesc) {
    const uint32_t A0 = smem_u32(&s->a[0][0][0]), PA = AF * 16;      // a planes: h = 0, l1 = 1, l2 = 2
    const uint32_t B0 = smem_u32(&s->b[0][0][0]), PB = BF * 16;      // b planes: l2 = 0, l1 = 1, h = 2
    const uint32_t Z = smem_u32(&s->zero[0][0]);
    const uint32_t aH = A0, aL1 = A0 + PA, aL2 = A0 + 2 * PA, bL2 = B0, bL1 = B0 + PB, bH = B0 + 2 * PB;
    auto same = [&](uint32_t acc, uint32_t pa, uint32_t pb, bool first) {      // one plane pair, 9 taps: 4 double steps + 1 with a zero chunk
        for (int t = 0; t < 8; t += 2)
            mma_i8(tmem + acc * N, make_desc(pa + 16 * t, 16, 128), make_desc(pb + 16 * t, 16, 128), idesc, !(first && t == 0));
        mma_i8(tmem + acc * N, make_desc(pa + 16 * 8, 16, 128), make_desc(pb + 16 * 8, Z - (pb + 16 * 8), 128), idesc, 1);
    };
    auto cross = [&](uint32_t acc, uint32_t pa0, uint32_t pa1, uint32_t pb0, uint32_t pb1, bool first) {   // pa0.pb0 + pa1.pb1 per tap
        for (int t = 0; t < TAPS; ++t)
            mma_i8(tmem + acc * N, make_desc(pa0 + 16 * t, pa1 - pa0, 128), make_desc(pb0 + 16 * t, pb1 - pb0, 128), idesc, !(first && t == 0));
    };
    same(0, aH, bH, true);
    cross(1, aH, aL1, bL1, bH, true);            // h.l1' + l1.h'
    same(2, aL1, bL1, true);
    cross(2, aH, aL2, bL2, bH, false);           // h.l2' + l2.h'
}

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t *__restrict__ ga, const uint8_t *__restrict__ gb, int *__restrict__ out,
                                                    int reps, long long *__restrict__ cycles, int mode) {
    __shared__ Smem s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (mode == 3) mode = (blockIdx.x >= gridDim.x / 2) ? 2 : 1;   // concurrency test: first half of the grid issues MMAs, second half reads TMEM
    for (int i = tid; i < NPL * AF * 16; i += 128) (&s.a[0][0][0])[i] = ga[i];
    for (int i = tid; i < NPL * BF * 16; i += 128) (&s.b[0][0][0])[i] = gb[i];
    for (int i = tid; i < N * 16; i += 128) (&s.zero[0][0])[i] = 0;
    if (tid == 0) {
        mbar_init(&s.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores above -> async-proxy reads of the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    const uint32_t idesc = idesc_i8(N);
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (mode != 2) {
            if (tid == 0) {
                issue_block(&s, tmem, idesc);
                mma_commit(&s.bar);
            }
            mbar_wait(&s.bar, phase);
            phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        int acc = 0;
        if (mode != 1)
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < N; c += 16) {
                int v[16];
                tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + a * N + c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (r == reps - 1) {
                    for (int i = 0; i < 16; ++i) out[(a * M + tid) * N + c + i] = v[i];
                } else {
                    for (int i = 0; i < 16; ++i) acc += v[i];
                }
            }
        if (acc == 0x7fffffff) out[0] = acc;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<uint8_t> a(NPL * AF * 16), b(NPL * BF * 16);
    srand(7);
    for (auto &x : a) x = rand() & 255;
    for (auto &x : b) x = rand() & 255;
    // CPU: limbs of row m, tap t, bin k: a[plane][m + t][k] (k < 16: the pad bytes take part, as in the hardware)
    std::vector<int> want(3 * M * N);
    auto A = [&](int pl, int f, int k) { return (int)a[(pl * AF + f) * 16 + k]; };
    auto B = [&](int pl, int f, int k) { return (int)b[(pl * BF + f) * 16 + k]; };   // planes stored l2, l1, h
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            long long s0 = 0, s1 = 0, s2 = 0;
            for (int t = 0; t < TAPS; ++t)
                for (int k = 0; k < 16; ++k) {
                    const int ah = A(0, m + t, k), al1 = A(1, m + t, k), al2 = A(2, m + t, k);
                    const int bh = B(2, n + t, k), bl1 = B(1, n + t, k), bl2 = B(0, n + t, k);
                    s0 += ah * bh;
                    s1 += ah * bl1 + al1 * bh;
                    s2 += al1 * bl1 + ah * bl2 + al2 * bh;
                }
            want[(0 * M + m) * N + n] = (int)s0; want[(1 * M + m) * N + n] = (int)s1; want[(2 * M + m) * N + n] = (int)s2;
        }
    uint8_t *da, *db; int *dout; long long *dc;
    cudaMalloc(&da, a.size()); cudaMalloc(&db, b.size()); cudaMalloc(&dout, want.size() * 4); cudaMalloc(&dc, 1024 * 8);
    cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xff, want.size() * 4);
    probe_kernel<<<1, 128>>>(da, db, dout, 1, dc, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<int> got(want.size());
    cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    for (int acc = 0; acc < 3; ++acc) {
        long bad = 0; int fm = -1, fn = -1;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n)
                if (got[(acc * M + m) * N + n] != want[(acc * M + m) * N + n]) { if (!bad) { fm = m; fn = n; } ++bad; }
        printf("acc%d: %ld mismatches of %d", acc, bad, M * N);
        if (bad) printf("  first (m=%d, n=%d): got %d want %d", fm, fn, got[(acc * M + fm) * N + fn], want[(acc * M + fm) * N + fn]);
        printf("\n");
    }
    // timing: one CTA alone, then one CTA per SM, then 4 CTAs per SM (TMEM: 4 x 128... this probe allocates 256 => 2 per SM)
    for (int mode = 0; mode < 4; ++mode)
    for (int grid : {1, 148, 296}) {
        const int reps = 200;
        probe_kernel<<<grid, 128>>>(da, db, dout, reps, dc, mode);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("timing launch: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> cyc(grid);
        cudaMemcpy(cyc.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
        if (mode == 3 && grid == 296) {
            long long m1 = 0, m2 = 0;
            for (int i = 0; i < 148; ++i) { m1 = cyc[i] > m1 ? cyc[i] : m1; m2 = cyc[148 + i] > m2 ? cyc[148 + i] : m2; }
            printf("mode 3 grid 296: MMA-only CTAs %.0f cycles per block, ld-only CTAs %.0f cycles per block (sharing SMs)\n", (double)m1 / reps, (double)m2 / reps);
        }
        printf("mode %d (0 full, 1 MMA only, 2 ld only) grid %d: %.0f cycles per 128x%d block, %.2f cells/clk/CTA\n", mode, grid, (double)mx / reps, N,
               (double)M * N * reps / mx);
    }
    return 0;
}
