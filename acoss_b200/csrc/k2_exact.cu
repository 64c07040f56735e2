// K2 (exact path) — reference-order cross-similarity, percentile thresholds and bit-packed CRP.
//
// Restates essentia ChromaCrossSimilarity (call site
// /root/reference/acoss/algorithms/rqa_serra09.py:60-66) with the SAME floating-point operation
// order as the oracle, so distances, thresholds and CRP bits are bit-identical by construction:
//   dot  : float32 products, sequential float64 accumulation over (t, bin), narrowed to float32
//   item : f32(f32(aa - 2ab) + bb),  d = sqrtf(item)
//   thr  : essentia percentile() (no floor==ceil guard unless integer_guard), exact order statistics
//   crp  : (thrQ[i] - d >= 0) && (thrR[j] - d >= 0)
// This path materialises a per-slot float32 distance tile in scratch HBM; it is the debugging /
// fallback path (ACOSS_CRP_EXACT, or pairs the fast path flags).  The production path
// (k2_fast.cu) never writes float distances to HBM.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// prep: rotate reference frames by oti, exact a.a / b.b per stacked window
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) exact_prep_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                         const int32_t *__restrict__ oti, int64_t first,
                                                         const int32_t *__restrict__ slot_map, int m, int tau,
                                                         int incr, int f32acc, int max_frames, int max_rows, int max_cols,
                                                         float *__restrict__ rrot_all, float *__restrict__ aa_all,
                                                         float *__restrict__ bb_all) {
    const int slot = blockIdx.x;
    const int64_t k = slot_map ? slot_map[slot] : first + slot;
    if (k < 0) return;                                        // unused entry of a fallback map
    const int q = pairs[2 * k], r = pairs[2 * k + 1], s = oti[k] % NBINS;
    const int nq = (int)(ts.offsets[q + 1] - ts.offsets[q]), nr = (int)(ts.offsets[r + 1] - ts.offsets[r]);
    const int M = nq - incr, N = nr - incr;
    const float *Q = ts.frames + ts.offsets[q] * NBINS;
    const float *R = ts.frames + ts.offsets[r] * NBINS;
    float *rrot = rrot_all + (int64_t)slot * max_frames * NBINS;
    for (int idx = threadIdx.x; idx < nr * NBINS; idx += blockDim.x) {
        const int f = idx / NBINS, b = idx - f * NBINS;
        rrot[idx] = R[f * NBINS + rot_src(b, s)];
    }
    __syncthreads();
    float *aa = aa_all + (int64_t)slot * max_rows, *bb = bb_all + (int64_t)slot * max_cols;
    for (int i = threadIdx.x; i < M + N; i += blockDim.x) {
        const bool isq = i < M;
        const float *src = isq ? Q + (int64_t)i * NBINS : rrot + (int64_t)(i - M) * NBINS;
        double acc = 0.0;
        for (int t = 0; t < m; ++t) {
            const float *fr = src + (int64_t)t * tau * NBINS;
#pragma unroll
            for (int b = 0; b < NBINS; ++b) acc = f32acc ? acc_f32prod_f32(acc, fr[b], fr[b]) : acc_f32prod(acc, fr[b], fr[b]);
        }
        if (isq) aa[i] = (float)acc; else bb[i - M] = (float)acc;
    }
}

// ------------------------------------------------------------------------------------------------
// distances: CTA tile TR rows x TC columns, frames staged in shared memory (pitch 13 -> no bank
// conflicts), each thread owns RPT rows x CPT columns with float64 accumulators
// ------------------------------------------------------------------------------------------------
constexpr int TR = 16, TC = 128, RPT = 2, CPT = 4, SP = 13;

__global__ void __launch_bounds__(256) exact_dist_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                         int64_t first, const int32_t *__restrict__ slot_map,
                                                         int m, int tau, int incr, int f32acc, int max_frames, int max_rows,
                                                         int max_cols, const float *__restrict__ rrot_all,
                                                         const float *__restrict__ aa_all,
                                                         const float *__restrict__ bb_all, float *__restrict__ D_all,
                                                         int64_t ldd, uint32_t *__restrict__ status) {
    extern __shared__ float smem[];
    const int slot = blockIdx.z;
    const int64_t k = slot_map ? slot_map[slot] : first + slot;
    if (k < 0) return;
    const int q = pairs[2 * k], r = pairs[2 * k + 1];
    const int nq = (int)(ts.offsets[q + 1] - ts.offsets[q]), nr = (int)(ts.offsets[r + 1] - ts.offsets[r]);
    const int M = nq - incr, N = nr - incr;
    const int i0 = blockIdx.y * TR, j0 = blockIdx.x * TC;
    if (i0 >= M || j0 >= N) return;
    const int span = (m - 1) * tau;
    const int qrows = TR + span, rrows = TC + span;
    float *sq = smem, *sr = smem + qrows * SP;
    const float *Q = ts.frames + ts.offsets[q] * NBINS;
    const float *R = rrot_all + (int64_t)slot * max_frames * NBINS;
    for (int idx = threadIdx.x; idx < qrows * NBINS; idx += blockDim.x) {
        const int f = idx / NBINS, b = idx - f * NBINS;
        sq[f * SP + b] = (i0 + f < nq) ? Q[(int64_t)(i0 + f) * NBINS + b] : 0.f;
    }
    for (int idx = threadIdx.x; idx < rrows * NBINS; idx += blockDim.x) {
        const int f = idx / NBINS, b = idx - f * NBINS;
        sr[f * SP + b] = (j0 + f < nr) ? R[(int64_t)(j0 + f) * NBINS + b] : 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[RPT][CPT];
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[a][c] = 0.0;
    for (int t = 0; t < m; ++t) {
#pragma unroll
        for (int b = 0; b < NBINS; ++b) {
            float qv[RPT], rv[CPT];
#pragma unroll
            for (int a = 0; a < RPT; ++a) qv[a] = sq[(warp * RPT + a + t * tau) * SP + b];
#pragma unroll
            for (int c = 0; c < CPT; ++c) rv[c] = sr[(lane + 32 * c + t * tau) * SP + b];
#pragma unroll
            for (int a = 0; a < RPT; ++a)
#pragma unroll
                for (int c = 0; c < CPT; ++c)
                    acc[a][c] = f32acc ? acc_f32prod_f32(acc[a][c], qv[a], rv[c]) : acc_f32prod(acc[a][c], qv[a], rv[c]);
        }
    }
    const float *aa = aa_all + (int64_t)slot * max_rows, *bb = bb_all + (int64_t)slot * max_cols;
    float *D = D_all + (int64_t)slot * max_rows * ldd;
    bool nan = false;
#pragma unroll
    for (int a = 0; a < RPT; ++a) {
        const int i = i0 + warp * RPT + a;
        if (i >= M) continue;
        const float av = aa[i];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int j = j0 + lane + 32 * c;
            if (j >= N) continue;
            const float ab = (float)acc[a][c];
            const float item = __fadd_rn(__fsub_rn(av, __fmul_rn(2.f, ab)), bb[j]);
            const float d = __fsqrt_rn(item);
            nan |= (d != d);
            D[(int64_t)i * ldd + j] = d;
        }
    }
    if (nan) atomicOr(&status[k], PAIR_ST_NAN);
}

// ------------------------------------------------------------------------------------------------
// percentile thresholds: exact k-th / (k+1)-th order statistic by 4-pass 8-bit radix select on the
// float32 bit pattern (d >= 0 so the unsigned order is the float order), then essentia's
// interpolation formula in float32.  One CTA per row (COLS=false) or per column (COLS=true).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float essentia_percentile(float s_fk, float s_ck, float kf, float fk, float ck,
                                                     int guard) {
    if (guard && fk == ck) return s_fk;
    const float d0 = __fmul_rn(s_fk, __fsub_rn(ck, kf));
    const float d1 = __fmul_rn(s_ck, __fsub_rn(kf, fk));
    return __fadd_rn(d0, d1);
}

template <bool COLS>
__global__ void __launch_bounds__(256) exact_select_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                           int64_t first, const int32_t *__restrict__ slot_map,
                                                           int incr, int max_rows, int max_cols,
                                                           const float *__restrict__ D_all, int64_t ldd, float qperc,
                                                           int guard, int64_t out_base, float *__restrict__ thr_all) {
    extern __shared__ uint32_t keys[];
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_k, s_cnt_le, s_min_gt;
    const int slot = blockIdx.y;
    const int64_t k = slot_map ? slot_map[slot] : first + slot;
    if (k < 0) return;
    const int64_t oslot = out_base < 0 ? slot : k - out_base;  // out_base < 0: outputs go to the scratch slot itself
    const int q = pairs[2 * k], r = pairs[2 * k + 1];
    const int M = (int)(ts.offsets[q + 1] - ts.offsets[q]) - incr, N = (int)(ts.offsets[r + 1] - ts.offsets[r]) - incr;
    const int line = blockIdx.x;
    const int nlines = COLS ? N : M, L = COLS ? M : N;
    if (line >= nlines) return;
    const float *D = D_all + (int64_t)slot * max_rows * ldd;
    const float *base = COLS ? D + line : D + (int64_t)line * ldd;
    const int64_t stride = COLS ? ldd : 1;
    for (int e = threadIdx.x; e < L; e += blockDim.x) keys[e] = __float_as_uint(base[(int64_t)e * stride]);
    const float kf = (L > 1) ? __fmul_rn((float)(L - 1), qperc) : __fmul_rn((float)L, qperc);
    const float fk = floorf(kf), ck = ceilf(kf);
    const int ifk = (int)fk, ick = min((int)ck, L - 1);
    if (threadIdx.x == 0) { s_prefix = 0u; s_k = (unsigned)ifk; }
    unsigned mask = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix;
        for (int e = threadIdx.x; e < L; e += blockDim.x) {
            const unsigned key = keys[e];
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned cum = 0, kk = s_k;
            for (int b = 0; b < 256; ++b) {
                const unsigned h = hist[b];
                if (cum + h > kk) { s_prefix = prefix | ((unsigned)b << shift); s_k = kk - cum; break; }
                cum += h;
            }
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    const unsigned vfk = s_prefix;
    if (threadIdx.x == 0) { s_cnt_le = 0u; s_min_gt = 0xffffffffu; }
    __syncthreads();
    unsigned cnt = 0, mn = 0xffffffffu;
    for (int e = threadIdx.x; e < L; e += blockDim.x) {
        const unsigned key = keys[e];
        if (key <= vfk) ++cnt; else mn = min(mn, key);
    }
    atomicAdd(&s_cnt_le, cnt);
    atomicMin(&s_min_gt, mn);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned vck = ((int)s_cnt_le >= ick + 1) ? vfk : s_min_gt;
        const float thr = essentia_percentile(__uint_as_float(vfk), __uint_as_float(vck), kf, fk, ck, guard);
        thr_all[oslot * (COLS ? max_cols : max_rows) + line] = thr;
    }
}

// ------------------------------------------------------------------------------------------------
// emit: one warp per 32 cells -> one ballot -> one CRP word.  Pad words/bits are written as zero.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) exact_emit_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                         int64_t first, const int32_t *__restrict__ slot_map,
                                                         int incr, int max_rows, int max_cols,
                                                         const float *__restrict__ D_all, int64_t ldd,
                                                         const float *__restrict__ thr_q_all,
                                                         const float *__restrict__ thr_r_all, int words,
                                                         int64_t crp_words, int64_t out_base, int strict,
                                                         uint32_t *__restrict__ crp_all) {
    const int slot = blockIdx.z;
    const int64_t k = slot_map ? slot_map[slot] : first + slot;
    if (k < 0) return;
    const int64_t oslot = out_base < 0 ? slot : k - out_base;
    const int q = pairs[2 * k], r = pairs[2 * k + 1];
    const int M = (int)(ts.offsets[q + 1] - ts.offsets[q]) - incr, N = (int)(ts.offsets[r + 1] - ts.offsets[r]) - incr;
    const int i = blockIdx.y;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= M || w >= words) return;
    const int j = w * 32 + lane;
    bool bit = false;
    if (j < N) {
        const float d = D_all[((int64_t)slot * max_rows + i) * ldd + j];
        const float tq = thr_q_all[oslot * max_rows + i], tr = thr_r_all[oslot * max_cols + j];
        const float xq = __fsub_rn(tq, d), xr = __fsub_rn(tr, d);        // heaviside arguments (F2)
        bit = strict ? (xq > 0.f && xr > 0.f) : (xq >= 0.f && xr >= 0.f);
    }
    const unsigned word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) crp_all[oslot * crp_words + (int64_t)i * words + w] = word;
}

int launch_k2_exact(const TrackSet &ts, const int32_t *pairs, const int32_t *oti, int64_t first, int n,
                    const acoss_params &p, const SlotGeom &g, const ExactScratch &sc, uint32_t *crp,
                    float *thr_q, float *thr_r, uint32_t *status, const int32_t *slot_map, cudaStream_t st,
                    int64_t *launches, bool out_by_slot) {
    // scratch slot = blockIdx; pair index k = slot_map ? slot_map[slot] : first + slot (negative map entries are
    // skipped); outputs (CRP, thresholds) land in chunk slot k - first, so a fallback re-run overwrites the pair's
    // own slot — or, with out_by_slot, in the scratch slot itself (deferred fallback rounds).
    const int64_t out_base = out_by_slot ? -1 : first;
    if (n <= 0) return ACOSS_OK;
    const int incr = p.f4_keep_last ? (p.m - 1) * p.tau : p.m * p.tau;   // F4
    const float qperc = (float)((double)(p.kappa * 100.f) / 100.);   // App. A4 float32 round trip
    exact_prep_kernel<<<n, 256, 0, st>>>(ts, pairs, oti, first, slot_map, p.m, p.tau, incr, p.f3_float_acc, ts.max_frames,
                                         g.max_rows, g.max_cols, sc.rrot, sc.aa, sc.bb);
    CUDA_TRY(cudaGetLastError());
    const int span = (p.m - 1) * p.tau;
    const size_t smem_d = (size_t)(TR + span + TC + span) * SP * sizeof(float);
    if (smem_d > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(exact_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
    dim3 gd((g.max_cols + TC - 1) / TC, (g.max_rows + TR - 1) / TR, n);
    exact_dist_kernel<<<gd, 256, smem_d, st>>>(ts, pairs, first, slot_map, p.m, p.tau, incr, p.f3_float_acc, ts.max_frames,
                                               g.max_rows, g.max_cols, sc.rrot, sc.aa, sc.bb, sc.D, sc.ldd, status);
    CUDA_TRY(cudaGetLastError());
    const size_t smem_r = (size_t)g.max_cols * 4, smem_c = (size_t)g.max_rows * 4;
    if (smem_r > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(exact_select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    if (smem_c > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(exact_select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
    exact_select_kernel<false><<<dim3(g.max_rows, n), 256, smem_r, st>>>(ts, pairs, first, slot_map, incr, g.max_rows,
                                                                        g.max_cols, sc.D, sc.ldd, qperc,
                                                                        p.integer_guard, out_base, thr_q);
    CUDA_TRY(cudaGetLastError());
    exact_select_kernel<true><<<dim3(g.max_cols, n), 256, smem_c, st>>>(ts, pairs, first, slot_map, incr, g.max_rows,
                                                                       g.max_cols, sc.D, sc.ldd, qperc,
                                                                       p.integer_guard, out_base, thr_r);
    CUDA_TRY(cudaGetLastError());
    dim3 ge((g.words + 7) / 8, g.max_rows, n);
    exact_emit_kernel<<<ge, 256, 0, st>>>(ts, pairs, first, slot_map, incr, g.max_rows, g.max_cols, sc.D, sc.ldd,
                                          thr_q, thr_r, g.words, g.crp_words, out_base, p.f2_strict, crp);
    CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 5;
    return ACOSS_OK;
}
