// EarlyFusion pair scoring on the device: cross-similarity matrices of the block features and the
// "early" similarity-network fusion of one pair.
//
// Replaces, per pair, the body of EarlyFusion.similarity
// (/root/reference/acoss/algorithms/earlyfusion_traile.py:157-198):
//   get_csm(X, Y)                          utils/cross_recurrence.py:31-48     (mfccs, ssms)
//   get_csm_blocked_oti(.., get_csm_cosine) utils/cross_recurrence.py:54-134   (chromas)
//   getWCSM(CSM, K, K)                     utils/similarity_fusion.py:38-54
//   exp(-sum of the three WCSMs)           earlyfusion_traile.py:178-182
// The binarisation (k4_knn.cu) and the Smith-Waterman DP (k3_dp.cu) consume the float64 matrices
// straight from HBM; nothing goes back to the host between the stages.
//
// Arithmetic: float64 throughout (the reference's numba/numpy code computes in the dtype of its
// inputs; its float32 features make BLAS-order-dependent float32 CSMs — the float64 value is the
// one every BLAS approximates, and it is what the pinned oracle computes).  These are genuine
// GEMMs (K = 480 / 1000 / 1225) but float64 ones, so tcgen05 (no float64 kind) does not apply; the
// float64 tensor path of sm_100a is the warp-level DMMA.  Three generations of the contraction live
// here, selectable with ACOSS_EF_CSM=1|2|3 for A/B timing (profiles/r1_ef.md):
//   1  ef_csm_kernel   DFMA, 8x4 register tiles, k-major shared tiles          13.7 TFLOP/s
//   2  ef_csm2_kernel  DFMA, 8x8 register tiles, row-major cp.async tiles      19.6 TFLOP/s
//   3  ef_csm3_kernel  DMMA m8n8k4, 2x2 warps of 32x32, cp.async (production)   30.4 TFLOP/s
// The DFMA kernels add the k terms of a cell in ascending k, one fused multiply-add per term; the DMMA kernel
// produces bit-identical matrices (tools/ef_cmp_generations.py on a B200: all four matrices of a 300-block pair equal
// across the three generations), i.e. the instruction behaves as the same ascending-k FMA chain.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// feature upload: [rows][d] (float32 or float64) -> [rows][dp] float64, zero padded to dp = 16 * ceil(d / 16)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void ef_widen_kernel(const T *__restrict__ src, int64_t rows, int d, int dp, double *__restrict__ dst) {
    const int64_t total = rows * dp;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / dp;
        const int k = (int)(e - r * dp);
        dst[e] = k < d ? (double)src[r * d + k] : 0.0;
    }
}

// One warp per block row.  mode 0: sq[row] = sum x^2 (np.sum(X**2, 1), cross_recurrence.py:45).
// mode 1: row /= sqrt(sum x^2), zero norms replaced by 1 (cross_recurrence.py:67-71; the norm of a row is
// the same for every roll of its chroma groups, so normalising before the roll gives the same quotients).
__global__ void __launch_bounds__(256) ef_rownorm_kernel(double *__restrict__ feat, int64_t rows, int dp, int mode,
                                                         double *__restrict__ sq) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    double *x = feat + row * dp;
    double s = 0.0;
    for (int k = lane; k < dp; k += 32) s = fma(x[k], x[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (mode == 0) {
        if (lane == 0) sq[row] = s;
    } else {
        double nrm = sqrt(s);
        if (nrm == 0.0) nrm = 1.0;
        for (int k = lane; k < dp; k += 32) x[k] = x[k] / nrm;
    }
}

int launch_ef_widen(const void *src, int elem_size, int64_t rows, int d, int dp, double *dst, cudaStream_t st) {
    if (rows <= 0) return ACOSS_OK;
    const int64_t total = rows * dp;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    if (elem_size == 4) ef_widen_kernel<float><<<blocks, 256, 0, st>>>((const float *)src, rows, d, dp, dst);
    else ef_widen_kernel<double><<<blocks, 256, 0, st>>>((const double *)src, rows, d, dp, dst);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

int launch_ef_rownorm(double *feat, int64_t rows, int dp, int mode, double *sq, cudaStream_t st) {
    if (rows <= 0) return ACOSS_OK;
    ef_rownorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(feat, rows, dp, mode, sq);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

// ---------------------------------------------------------------------------------------------
// OTI of the song-level chroma medians: argmax_i sum(roll(C1, i) * C2), first maximum
// (cross_recurrence.py:94-103).  The 12 products are added in the order numpy's pairwise sum uses
// for a contiguous 12-element array (8 running sums folded as a tree, then the 4 leftovers).
// ---------------------------------------------------------------------------------------------
__global__ void ef_oti_kernel(const double *__restrict__ cmed, const int32_t *__restrict__ pairs, int n,
                              int32_t *__restrict__ oti) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double *c1 = cmed + (int64_t)pairs[2 * k] * NBINS;
    const double *c2 = cmed + (int64_t)pairs[2 * k + 1] * NBINS;
    double a[NBINS], b[NBINS];
#pragma unroll
    for (int t = 0; t < NBINS; ++t) { a[t] = c1[t]; b[t] = c2[t]; }
    int best = 0;
    double bestv = 0.0;
    for (int s = 0; s < NBINS; ++s) {
        double p[NBINS];
#pragma unroll
        for (int t = 0; t < NBINS; ++t) {
            int src = t - s;
            if (src < 0) src += NBINS;                       // np.roll(C1, s)[t] = C1[(t - s) mod 12]
            double av = a[0];
#pragma unroll
            for (int u = 1; u < NBINS; ++u) av = (u == src) ? a[u] : av;
            p[t] = __dmul_rn(av, b[t]);
        }
        double r = __dadd_rn(__dadd_rn(__dadd_rn(p[0], p[1]), __dadd_rn(p[2], p[3])),
                             __dadd_rn(__dadd_rn(p[4], p[5]), __dadd_rn(p[6], p[7])));
        r = __dadd_rn(r, p[8]); r = __dadd_rn(r, p[9]); r = __dadd_rn(r, p[10]); r = __dadd_rn(r, p[11]);
        if (s == 0 || r > bestv) { bestv = r; best = s; }
    }
    oti[k] = best;
}

int launch_ef_oti(const double *cmed, const int32_t *pairs, int n, int32_t *oti, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    ef_oti_kernel<<<(n + 127) / 128, 128, 0, st>>>(cmed, pairs, n, oti);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

// ---------------------------------------------------------------------------------------------
// CSM: one 64x64 tile of one pair per CTA
// ---------------------------------------------------------------------------------------------
#define EF_BM 64
#define EF_BN 64
#define EF_BK 16
#define EF_PITCH 66     // doubles per k-row of a shared tile (16-byte aligned rows, staggered banks)

// MODE 0: sqrt(max(0, (|x|^2 + |y|^2) - 2 x.y)); MODE 1: 1 - xhat.yhat with the 12-bin groups of x rolled by oti
template <int MODE>
__global__ void __launch_bounds__(128) ef_csm_kernel(const double *__restrict__ feat, int dp, int d,
                                                     const double *__restrict__ sq,
                                                     const int64_t *__restrict__ offsets,
                                                     const int32_t *__restrict__ pairs,
                                                     const int32_t *__restrict__ oti_a,
                                                     double *__restrict__ csm, int64_t slot_elems, int tiles_n) {
    __shared__ __align__(16) double As[EF_BK][EF_PITCH];
    __shared__ __align__(16) double Bs[EF_BK][EF_PITCH];
    const int slot = blockIdx.y;
    const int q = pairs[2 * slot], r = pairs[2 * slot + 1];
    const int64_t oq = offsets[q], orr = offsets[r];
    const int M = (int)(offsets[q + 1] - oq), N = (int)(offsets[r + 1] - orr);
    const int m0 = (blockIdx.x / tiles_n) * EF_BM, n0 = (blockIdx.x % tiles_n) * EF_BN;
    if (m0 >= M || n0 >= N) return;
    const int tid = threadIdx.x;
    // global -> register staging: thread = (tile row, 8 consecutive k)
    const int lrow = tid & 63, kh = (tid >> 6) * 8;
    const double *ga = feat + (oq + min(m0 + lrow, M - 1)) * (int64_t)dp;
    const double *gb = feat + (orr + min(n0 + lrow, N - 1)) * (int64_t)dp;
    int rot = 0;
    if (MODE == 1) rot = oti_a[slot];
    // register tile: rows 2rg + {0,1} + 16 s (s = 0..3), columns 2cg + {0,1} + 32 h (h = 0,1)
    const int rg = tid >> 4, cg = tid & 15;
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    double ra[8], rb[8];
    auto fetch = [&](int k0) {
        if (MODE == 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = k0 + kh + e;
                int src = k;
                if (k < d) {                                 // np.roll(X1, oti, axis=2): out[b] = in[(b - oti) mod 12]
                    const int t = k / NBINS, b = k - t * NBINS;
                    src = t * NBINS + rot_src(b, rot);
                }
                ra[e] = ga[src];
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(ga + k0 + kh + e);
                ra[e] = v.x; ra[e + 1] = v.y;
            }
        }
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(gb + k0 + kh + e);
            rb[e] = v.x; rb[e + 1] = v.y;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < dp; k0 += EF_BK) {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; ++e) { As[kh + e][lrow] = ra[e]; Bs[kh + e][lrow] = rb[e]; }
        __syncthreads();
        if (k0 + EF_BK < dp) fetch(k0 + EF_BK);
#pragma unroll
        for (int kk = 0; kk < EF_BK; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double2 v = *reinterpret_cast<const double2 *>(&As[kk][2 * rg + 16 * s]);
                a[2 * s] = v.x; a[2 * s + 1] = v.y;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double2 v = *reinterpret_cast<const double2 *>(&Bs[kk][2 * cg + 32 * h]);
                b[2 * h] = v.x; b[2 * h + 1] = v.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
    double *out = csm + (int64_t)slot * slot_elems;
    double sy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = n0 + 2 * cg + (j & 1) + 32 * (j >> 1);
        sy[j] = (MODE == 0 && col < N) ? sq[orr + col] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + 2 * rg + (i & 1) + 16 * (i >> 1);
        if (row >= M) continue;
        const double sx = (MODE == 0) ? sq[oq + row] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + 2 * cg + (j & 1) + 32 * (j >> 1);
            if (col >= N) continue;
            double v;
            if (MODE == 0) {
                double c2 = __dsub_rn(__dadd_rn(sx, sy[j]), __dmul_rn(2.0, acc[i][j]));
                c2 = c2 < 0.0 ? 0.0 : c2;
                v = sqrt(c2);
            } else {
                v = __dsub_rn(1.0, acc[i][j]);
            }
            out[(int64_t)row * N + col] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CSM, second generation.  ncu on the kernel above (profiles/r1_ef_csm.md): the shared-memory pipe is the
// limiter (l1tex throughput 91 %, FP64 pipe 45 %) — an LDS.128 is served per half-warp, 128 B per wavefront,
// so the 2 x 16 thread layout pays 4 wavefronts per B fragment.  Here: 64-thread CTAs (2 warps), 8 x 8
// accumulators per thread, warp = 4 row groups x 8 column groups, tiles kept ROW-major in shared memory
// ([row][k], pitch 18 doubles) so that one LDS.128 delivers two consecutive k of a row and every half-warp
// touches exactly 128 distinct bytes (B) or 32 (A): 32 wavefronts per 128 DFMA per warp, half the pipe.
// Global -> shared goes through cp.async (16 B chunks, two-stage ring), no staging registers.  The k order
// of every accumulator is unchanged (k ascending, one FMA per k), so the results are bit-identical to the
// first-generation kernel's.
// ---------------------------------------------------------------------------------------------
#define EF2_LD 18
#define EF2_MAXDP 2048   // widest (padded) chroma block the rolled-column map covers
__device__ __forceinline__ void ef_cp_async16(void *smem, const void *g) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void ef_cp_async8(void *smem, const void *g) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(g) : "memory");
}

// 16 k of one 64 x 64 tile for one thread: acc[i][j] += A[row_i][k] * B[col_j][k], k ascending
#ifdef ACOSS_EF_GENERATIONS   // superseded generation 2 (DFMA 8x8, cp.async): compiled only for tools/ef_cmp_generations.py
template <bool FULL>
__device__ __forceinline__ void ef_tile_fma(const double *Ap, const double *Bp, int jmax, double (&acc)[8][8]) {
#pragma unroll
    for (int k2 = 0; k2 < EF_BK; k2 += 2) {
        double2 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const double2 *>(Ap + i * 4 * EF2_LD + k2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (!FULL && j >= jmax) break;                    // warp-uniform
            const double2 b = *reinterpret_cast<const double2 *>(Bp + j * 8 * EF2_LD + k2);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i][j] = fma(a[i].x, b.x, acc[i][j]);
                acc[i][j] = fma(a[i].y, b.y, acc[i][j]);
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(64) ef_csm2_kernel(const double *__restrict__ feat, int dp, int d,
                                                     const double *__restrict__ sq,
                                                     const int64_t *__restrict__ offsets,
                                                     const int32_t *__restrict__ pairs,
                                                     const int32_t *__restrict__ oti_a,
                                                     double *__restrict__ csm, int64_t slot_elems, int tiles_n) {
    __shared__ __align__(16) double As[2][EF_BM][EF2_LD];
    __shared__ __align__(16) double Bs[2][EF_BN][EF2_LD];
    __shared__ short kmap[MODE == 1 ? EF2_MAXDP : 1];
    const int slot = blockIdx.y;
    const int q = pairs[2 * slot], r = pairs[2 * slot + 1];
    const int64_t oq = offsets[q], orr = offsets[r];
    const int M = (int)(offsets[q + 1] - oq), N = (int)(offsets[r + 1] - orr);
    const int m0 = (blockIdx.x / tiles_n) * EF_BM, n0 = (blockIdx.x % tiles_n) * EF_BN;
    if (m0 >= M || n0 >= N) return;
    const int tid = threadIdx.x;
    const double *Abase = feat + oq * (int64_t)dp, *Bbase = feat + orr * (int64_t)dp;
    if (MODE == 1) {                                          // np.roll(X1, oti, axis=2): out[b] = in[(b - oti) mod 12]
        const int rot = oti_a[slot];
        for (int k = tid; k < dp; k += 64) {
            int src = k;
            if (k < d) {
                const int t = k / NBINS, b = k - t * NBINS;
                src = t * NBINS + rot_src(b, rot);
            }
            kmap[k] = (short)src;
        }
        __syncthreads();
    }
    auto issue = [&](int k0, int buf) {
        const int ch = tid & 7;
        if (MODE == 1) {
#pragma unroll 1
            for (int e = 0; e < 16; ++e) {                    // rolled chroma groups: element-wise gather
                const int id = tid + 64 * e, row = id >> 4, kk = id & 15;
                ef_cp_async8(&As[buf][row][kk], Abase + min(m0 + row, M - 1) * dp + (int)kmap[k0 + kk]);
            }
        }
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            const int row = (tid >> 3) + 8 * c;
            if (MODE == 0) ef_cp_async16(&As[buf][row][2 * ch], Abase + min(m0 + row, M - 1) * dp + k0 + 2 * ch);
            ef_cp_async16(&Bs[buf][row][2 * ch], Bbase + min(n0 + row, N - 1) * dp + k0 + 2 * ch);
        }
    };
    // register tile: rows 32 w + rg + 4 i (i = 0..7), columns cg + 8 j (j = 0..7)
    const int w = tid >> 5, rg = (tid >> 3) & 3, cg = tid & 7;
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    // ragged edge tiles: a warp whose 32 rows lie past M only helps loading; 8-column groups past N are skipped
    const bool warp_live = m0 + 32 * w < M;
    const int jmax = min(8, (N - n0 + 7) >> 3);
    const int nt = dp / EF_BK;
    issue(0, 0);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) issue((t + 1) * EF_BK, (t + 1) & 1);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncthreads();
        const double *Ap = &As[t & 1][32 * w + rg][0];
        const double *Bp = &Bs[t & 1][cg][0];
        if (warp_live) {
            if (jmax == 8) ef_tile_fma<true>(Ap, Bp, 8, acc);    // interior tile: fully unrolled, fragments prefetched
            else ef_tile_fma<false>(Ap, Bp, jmax, acc);          // ragged right edge: 8-column groups past N skipped
        }
        __syncthreads();
    }
    double *out = csm + (int64_t)slot * slot_elems;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + 32 * w + rg + 4 * i;
        if (row >= M) continue;
        const double sx = (MODE == 0) ? sq[oq + row] : 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + cg + 8 * j;
            if (col >= N) continue;
            double v;
            if (MODE == 0) {
                double c2 = __dsub_rn(__dadd_rn(sx, sq[orr + col]), __dmul_rn(2.0, acc[i][j]));
                c2 = c2 < 0.0 ? 0.0 : c2;
                v = sqrt(c2);
            } else {
                v = __dsub_rn(1.0, acc[i][j]);
            }
            out[(int64_t)row * N + col] = v;
        }
    }
}

#endif  // ACOSS_EF_GENERATIONS
// ---------------------------------------------------------------------------------------------
// CSM, third generation: the float64 tensor path (DMMA, mma.sync.m8n8k4.f64 — tcgen05 has no float64
// kind; on sm_100a the legacy warp-level MMA is the only float64 tensor instruction).  ncu on the second
// generation (profiles/r1_ef_csm.md): FP64 pipe 56 %, two warps per scheduler at 255 registers, stalls are
// the half-rate DFMA issue itself.  One DMMA replaces 8 DFMA per lane and needs one A and one B double
// per lane: 64 x 64 CTA tile, 2 x 2 warps of 32 x 32 (4 x 4 MMA tiles, 32 accumulator doubles per lane),
// 8 LDS.64 per 16 DMMA.  Shared tiles stay row-major [row][k] with a pitch of 20 doubles, so the 4 rows x
// 4 k a half-warp reads fall in 32 distinct banks.  cp.async two-stage ring as above.
// ---------------------------------------------------------------------------------------------
#define EF3_LD 20
__device__ __forceinline__ void ef_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 16 k of one 32 x 32 warp quadrant, IMAX x JMAX of its 4 x 4 MMA tiles (compile-time bounds: the skipped
// tiles are not even issued — a predicated-off DMMA still takes its tensor-pipe slot)
template <int IMAX, int JMAX>
__device__ __forceinline__ void ef_quad_mma(const double *Ap, const double *Bp, double (&acc)[4][4][2]) {
#pragma unroll
    for (int k4 = 0; k4 < EF_BK; k4 += 4) {
        double a[IMAX], b[JMAX];
#pragma unroll
        for (int i = 0; i < IMAX; ++i) a[i] = Ap[i * 8 * EF3_LD + k4];
#pragma unroll
        for (int j = 0; j < JMAX; ++j) b[j] = Bp[j * 8 * EF3_LD + k4];
#pragma unroll
        for (int i = 0; i < IMAX; ++i)
#pragma unroll
            for (int j = 0; j < JMAX; ++j) ef_dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}
template <int IMAX>
__device__ __forceinline__ void ef_quad_dispatch_j(int jmax, const double *Ap, const double *Bp, double (&acc)[4][4][2]) {
    switch (jmax) {
        case 4: ef_quad_mma<IMAX, 4>(Ap, Bp, acc); break;
        case 3: ef_quad_mma<IMAX, 3>(Ap, Bp, acc); break;
        case 2: ef_quad_mma<IMAX, 2>(Ap, Bp, acc); break;
        default: ef_quad_mma<IMAX, 1>(Ap, Bp, acc); break;
    }
}
__device__ __forceinline__ void ef_quad_dispatch(int imax, int jmax, const double *Ap, const double *Bp,
                                                 double (&acc)[4][4][2]) {
    switch (imax) {
        case 4: ef_quad_dispatch_j<4>(jmax, Ap, Bp, acc); break;
        case 3: ef_quad_dispatch_j<3>(jmax, Ap, Bp, acc); break;
        case 2: ef_quad_dispatch_j<2>(jmax, Ap, Bp, acc); break;
        default: ef_quad_dispatch_j<1>(jmax, Ap, Bp, acc); break;
    }
}

template <int MODE>
__global__ void __launch_bounds__(128) ef_csm3_kernel(const double *__restrict__ feat, int dp, int d,
                                                      const double *__restrict__ sq,
                                                      const int64_t *__restrict__ offsets,
                                                      const int32_t *__restrict__ pairs,
                                                      const int32_t *__restrict__ oti_a,
                                                      double *__restrict__ csm, int64_t slot_elems, int tiles_n) {
    __shared__ __align__(16) double As[2][EF_BM][EF3_LD];
    __shared__ __align__(16) double Bs[2][EF_BN][EF3_LD];
    __shared__ short kmap[MODE == 1 ? EF2_MAXDP : 1];
    const int slot = blockIdx.y;
    const int q = pairs[2 * slot], r = pairs[2 * slot + 1];
    const int64_t oq = offsets[q], orr = offsets[r];
    const int M = (int)(offsets[q + 1] - oq), N = (int)(offsets[r + 1] - orr);
    const int m0 = (blockIdx.x / tiles_n) * EF_BM, n0 = (blockIdx.x % tiles_n) * EF_BN;
    if (m0 >= M || n0 >= N) return;
    const int tid = threadIdx.x;
    const double *Abase = feat + oq * (int64_t)dp, *Bbase = feat + orr * (int64_t)dp;
    if (MODE == 1) {                                          // np.roll(X1, oti, axis=2): out[b] = in[(b - oti) mod 12]
        const int rot = oti_a[slot];
        for (int k = tid; k < dp; k += 128) {
            int src = k;
            if (k < d) {
                const int t = k / NBINS, b = k - t * NBINS;
                src = t * NBINS + rot_src(b, rot);
            }
            kmap[k] = (short)src;
        }
        __syncthreads();
    }
    auto issue = [&](int k0, int buf) {
        const int ch = tid & 7;
        if (MODE == 1) {
#pragma unroll 1
            for (int e = 0; e < 8; ++e) {                     // rolled chroma groups: element-wise gather
                const int id = tid + 128 * e, row = id >> 4, kk = id & 15;
                ef_cp_async8(&As[buf][row][kk], Abase + min(m0 + row, M - 1) * dp + (int)kmap[k0 + kk]);
            }
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            const int row = (tid >> 3) + 16 * c;
            if (MODE == 0) ef_cp_async16(&As[buf][row][2 * ch], Abase + min(m0 + row, M - 1) * dp + k0 + 2 * ch);
            ef_cp_async16(&Bs[buf][row][2 * ch], Bbase + min(n0 + row, N - 1) * dp + k0 + 2 * ch);
        }
    };
    // warp (wr, wc) owns rows 32 wr .. +31 and columns 32 wc .. +31; MMA tile (rt, ct) of it: lane holds
    // C[8 rt + g][8 ct + 2 tg + {0,1}], reads A[8 rt + g][k + tg] and B[k + tg][8 ct + g]  (g = lane / 4, tg = lane % 4)
    const int lane = tid & 31, wr = (tid >> 5) >> 1, wc = (tid >> 5) & 1, g = lane >> 2, tg = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const bool warp_live = m0 + 32 * wr < M && n0 + 32 * wc < N;      // ragged edge tiles: idle quadrants only load
    const int imax = min(4, (M - m0 - 32 * wr + 7) >> 3), jmax = min(4, (N - n0 - 32 * wc + 7) >> 3);
    const int nt = dp / EF_BK;
    issue(0, 0);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) issue((t + 1) * EF_BK, (t + 1) & 1);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncthreads();
        if (warp_live) {
            const double *Ap = &As[t & 1][32 * wr + g][tg];
            const double *Bp = &Bs[t & 1][32 * wc + g][tg];
            ef_quad_dispatch(imax, jmax, Ap, Bp, acc);            // ragged edges: only the MMA tiles inside M x N
        }
        __syncthreads();
    }
    if (!warp_live) return;
    double *out = csm + (int64_t)slot * slot_elems;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + 32 * wr + 8 * i + g;
        if (row >= M) continue;
        const double sx = (MODE == 0) ? sq[oq + row] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = n0 + 32 * wc + 8 * j + 2 * tg + h;
                if (col >= N) continue;
                double v;
                if (MODE == 0) {
                    double c2 = __dsub_rn(__dadd_rn(sx, sq[orr + col]), __dmul_rn(2.0, acc[i][j][h]));
                    c2 = c2 < 0.0 ? 0.0 : c2;
                    v = sqrt(c2);
                } else {
                    v = __dsub_rn(1.0, acc[i][j][h]);
                }
                out[(int64_t)row * N + col] = v;
            }
    }
}

static int ef_csm_generation() {
#ifdef ACOSS_EF_GENERATIONS
    static int gen = -1;
    if (gen < 0) {
        const char *e = getenv("ACOSS_EF_CSM");               // 1 / 2 select the DFMA kernels (A/B timing)
        gen = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 3;
    }
    return gen;
#else
    return 3;                                                 // the DMMA kernel; ef_csm_kernel only for chroma blocks wider than EF2_MAXDP
#endif
}

int launch_ef_csm(int mode, const double *feat, int dp, int d, const double *sq, const int64_t *offsets,
                  const int32_t *pairs, const int32_t *oti, int n, int max_rows, int max_cols, double *csm,
                  int64_t slot_elems, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    const int tiles_m = (max_rows + EF_BM - 1) / EF_BM, tiles_n = (max_cols + EF_BN - 1) / EF_BN;
    dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)n);
    if (ef_csm_generation() == 3 && (mode == 0 || dp <= EF2_MAXDP)) {
        static bool carve[16] = {false};                      // 40 KB static tiles per CTA: ask for the full carve-out
        int dev = 0;                                          // so that 5 CTAs fit an SM
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 16 && !carve[dev]) {
            cudaFuncSetAttribute(ef_csm3_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            cudaFuncSetAttribute(ef_csm3_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            carve[dev] = true;
        }
        if (mode == 0) ef_csm3_kernel<0><<<grid, 128, 0, st>>>(feat, dp, d, sq, offsets, pairs, oti, csm, slot_elems, tiles_n);
        else ef_csm3_kernel<1><<<grid, 128, 0, st>>>(feat, dp, d, sq, offsets, pairs, oti, csm, slot_elems, tiles_n);
        CUDA_TRY(cudaGetLastError());
        return ACOSS_OK;
    }
#ifdef ACOSS_EF_GENERATIONS
    if (ef_csm_generation() >= 2 && (mode == 0 || dp <= EF2_MAXDP)) {
        if (mode == 0) ef_csm2_kernel<0><<<grid, 64, 0, st>>>(feat, dp, d, sq, offsets, pairs, oti, csm, slot_elems, tiles_n);
        else ef_csm2_kernel<1><<<grid, 64, 0, st>>>(feat, dp, d, sq, offsets, pairs, oti, csm, slot_elems, tiles_n);
        CUDA_TRY(cudaGetLastError());
        return ACOSS_OK;
    }
#endif
    if (mode == 0) ef_csm_kernel<0><<<grid, 128, 0, st>>>(feat, dp, d, sq, offsets, pairs, oti, csm, slot_elems, tiles_n);
    else ef_csm_kernel<1><<<grid, 128, 0, st>>>(feat, dp, d, sq, offsets, pairs, oti, csm, slot_elems, tiles_n);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

// ---------------------------------------------------------------------------------------------
// getWCSM neighbourhood radii: mean of the K smallest entries of every row and of every column
// (similarity_fusion.py:48-52).  One thread per line keeps the KMAX smallest values seen so far in a
// sorted register array (max/min insertion chain, no data-dependent indexing); which of several equal
// values is kept does not change the sum.  Line l < M is row l, line M + j is column j.
// ---------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(128) ef_linestat_kernel(const double *__restrict__ csm, int64_t kind_stride,
                                                          int64_t slot_elems, const int64_t *__restrict__ offsets,
                                                          const int32_t *__restrict__ pairs, int K,
                                                          double *__restrict__ stat, int64_t stat_kind_stride,
                                                          int stat_pitch) {
    const int slot = blockIdx.y, kind = blockIdx.z;
    const int q = pairs[2 * slot], r = pairs[2 * slot + 1];
    const int M = (int)(offsets[q + 1] - offsets[q]), N = (int)(offsets[r + 1] - offsets[r]);
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= M + N) return;
    const double *base = csm + kind * kind_stride + (int64_t)slot * slot_elems;
    const double *p;
    int len, stride;
    if (l < M) { p = base + (int64_t)l * N; len = N; stride = 1; }
    else { p = base + (l - M); len = M; stride = N; }
    double a[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) a[t] = __longlong_as_double(0x7ff0000000000000ll);
    // eight loads in flight per thread (the insertion branch would otherwise serialise the memory latency)
    for (int e0 = 0; e0 < len; e0 += 8) {
        double v8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            v8[u] = (e0 + u < len) ? p[(int64_t)(e0 + u) * stride] : __longlong_as_double(0x7ff0000000000000ll);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double v = v8[u];
            if (v < a[KMAX - 1]) {
#pragma unroll
                for (int t = KMAX - 1; t > 0; --t) a[t] = fmax(a[t - 1], fmin(a[t], v));
                a[0] = fmin(a[0], v);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < KMAX; ++t)
        if (t < K) s = __dadd_rn(s, a[t]);
    stat[kind * stat_kind_stride + (int64_t)slot * stat_pitch + l] = s / (double)K;
}

int ef_linestat_max_k() { return 64; }

int launch_ef_linestat(const double *csm, int64_t kind_stride, int64_t slot_elems, const int64_t *offsets,
                       const int32_t *pairs, int n, int max_rows, int max_cols, int K, double *stat,
                       int64_t stat_kind_stride, int stat_pitch, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    dim3 grid((unsigned)((max_rows + max_cols + 127) / 128), (unsigned)n, 3);
    if (K <= 16)
        ef_linestat_kernel<16><<<grid, 128, 0, st>>>(csm, kind_stride, slot_elems, offsets, pairs, K, stat,
                                                     stat_kind_stride, stat_pitch);
    else
        ef_linestat_kernel<64><<<grid, 128, 0, st>>>(csm, kind_stride, slot_elems, offsets, pairs, K, stat,
                                                     stat_kind_stride, stat_pitch);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

// ---------------------------------------------------------------------------------------------
// early fusion: out = exp(-(W_mfccs + W_ssms + W_chromas)), W = exp(-d^2 / (2 (Mu Eps)^2)),
// Eps = ((rowmean + colmean) + d) / 3, Mu = 0.5   (similarity_fusion.py:52-54, earlyfusion_traile.py:178-182;
// same operation order, no contraction across the reference's separate numpy operations)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ef_fuse_kernel(const double *__restrict__ csm, int64_t kind_stride,
                                                      int64_t slot_elems, const int64_t *__restrict__ offsets,
                                                      const int32_t *__restrict__ pairs,
                                                      const double *__restrict__ stat, int64_t stat_kind_stride,
                                                      int stat_pitch, double mu, double *__restrict__ out) {
    const int slot = blockIdx.y;
    const int q = pairs[2 * slot], r = pairs[2 * slot + 1];
    const int M = (int)(offsets[q + 1] - offsets[q]), N = (int)(offsets[r + 1] - offsets[r]);
    const int64_t cells = (int64_t)M * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < cells; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / N), j = (int)(e - (int64_t)i * N);
        double total = 0.0;
#pragma unroll
        for (int kind = 0; kind < 3; ++kind) {
            const double dd = csm[kind * kind_stride + (int64_t)slot * slot_elems + e];
            const double *st = stat + kind * stat_kind_stride + (int64_t)slot * stat_pitch;
            const double eps = __ddiv_rn(__dadd_rn(__dadd_rn(st[i], st[M + j]), dd), 3.0);
            const double me = __dmul_rn(mu, eps);
            const double den = __dmul_rn(2.0, __dmul_rn(me, me));
            const double w = exp(__ddiv_rn(-__dmul_rn(dd, dd), den));
            total = __dadd_rn(total, w);
        }
        out[(int64_t)slot * slot_elems + e] = exp(-total);
    }
}

int launch_ef_fuse(const double *csm, int64_t kind_stride, int64_t slot_elems, const int64_t *offsets,
                   const int32_t *pairs, int n, int max_rows, int max_cols, const double *stat,
                   int64_t stat_kind_stride, int stat_pitch, double *out, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    const int64_t cells = (int64_t)max_rows * max_cols;
    dim3 grid((unsigned)std::min<int64_t>((cells + 1023) / 1024, 4096), (unsigned)n);
    ef_fuse_kernel<<<grid, 256, 0, st>>>(csm, kind_stride, slot_elems, offsets, pairs, stat, stat_kind_stride,
                                         stat_pitch, 0.5, out);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

// per-slot shapes, neighbour counts and CSM offsets for the k-NN kernel (cross_recurrence.py:151-155 rule:
// kappa == 0 -> all ones (-1), kappa < 1 -> int(np.round(kappa * columns)), else kappa)
__global__ void ef_geom_kernel(const int64_t *__restrict__ offsets, const int32_t *__restrict__ pairs, int n,
                               double kappa, int64_t slot_elems, int32_t *__restrict__ shapes,
                               int32_t *__restrict__ nn, int64_t *__restrict__ csm_off) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int q = pairs[2 * k], r = pairs[2 * k + 1];
    const int M = (int)(offsets[q + 1] - offsets[q]), N = (int)(offsets[r + 1] - offsets[r]);
    shapes[2 * k] = M;
    shapes[2 * k + 1] = N;
    int v;
    if (kappa == 0.0) v = -1;
    else if (kappa < 1.0) v = (int)rint(__dmul_rn(kappa, (double)N));      // np.round: half to even
    else v = (int)kappa;
    nn[k] = v;
    csm_off[k] = (int64_t)k * slot_elems;
}

int launch_ef_geom(const int64_t *offsets, const int32_t *pairs, int n, double kappa, int64_t slot_elems,
                   int32_t *shapes, int32_t *nn, int64_t *csm_off, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    ef_geom_kernel<<<(n + 127) / 128, 128, 0, st>>>(offsets, pairs, n, kappa, slot_elems, shapes, nn, csm_off);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}
