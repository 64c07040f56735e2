// K2 (fast path) — cross-similarity, per-row / per-column k-th-nearest selection and bit-packed CRP
// without ever writing a float distance matrix to HBM.
//
// Replaces essentia ChromaCrossSimilarity (call site
// /root/reference/acoss/algorithms/rqa_serra09.py:60-66).  Results are bit-identical to the
// reference-order arithmetic of k2_exact.cu; the structure is:
//
//  * The stacked squared distance is a 9-tap diagonal sum of the frame-level dot product
//    e[a][c] = <x_a, y_c> (12 FMAs):  item(i,j) = aa_i + bb_j - 2 * sum_t e[i+t][j+t].
//    A warp sweeps the streamed side (rows a) for a strip of owned columns c kept in REGISTERS
//    (RC columns per lane); the diagonal sum slides: T[a][c] = T[a-1][c-1] + e[a][c] - e[a-9][c-9]
//    (one warp shuffle for the lane boundary, one for the 9-rows-old value from the lane's
//    neighbour's register ring).  e is quantised once to fixed point (float magic-number add), so
//    T is an EXACT integer sliding sum: no drift, identical in every sweep and both orientations.
//  * These approximate items z (|z - exact| <= EPS units, bound in DESIGN.md §4.2) drive two
//    thread-private 64-bin histogram sweeps per orientation (columns: owned = reference; rows:
//    owned = query) that bracket the order statistics floor(k), ceil(k) of every row and column
//    to a handful of cells, and one emit sweep that writes every cell whose membership is certain
//    and appends the few uncertain cells (inside a bracket +- 2 EPS) to per-row / per-column
//    candidate lists.
//  * resolve kernels recompute ONLY the candidate cells in the reference's exact operation order
//    (float32 products, sequential float64 accumulation), pick the exact order statistics, apply
//    essentia's percentile formula and patch the candidate bits.  Consistency checks that make
//    the certain/uncertain split provably exact are evaluated per row/column; a pair that fails
//    one is flagged and re-run by k2_exact.cu.
#include <stdlib.h>

#include "k2_fast.cuh"

namespace {

constexpr int M9 = 9;               // frameStackSize handled by this path
constexpr int HALO = M9 - 1;        // owned columns a strip recomputes (8)
constexpr int NBIN = 64;            // histogram bins per level (+ one underflow and one overflow row)
constexpr int SPARSE_CAP = 512;     // live lines per pair and orientation the sparse refinement can take
constexpr int SBIN = 256;           // bins of the per-line sample histogram (select kernel)
constexpr int EPS = 128;            // bound on |z - exact item| in fixed-point units
constexpr int CAND_CAP = 128;       // candidates per row / column
constexpr int BRACKET_TARGET = 96;  // a bracket holding more cells than this is split by another histogram level
constexpr int WPC = 4;              // warps per CTA in the sweep kernels
constexpr int RCV = 4;              // owned frames per lane (register columns) in the sweep kernels

struct PairHdr {                    // per-slot header written by fast_prep_kernel
    int32_t nq, nr, Mx, Nx;         // frames and stacked windows of query / reference
    int32_t fk[2], ck[2];           // 0-based ranks floor(k), ceil(k): [0] rows (L = Nx), [1] columns (L = Mx)
    int32_t quirk[2];               // 1: threshold is forced to 0 (integer k without guard, F1)
    float kf[2];                    // fractional rank (float32, essentia arithmetic)
    int32_t lo1, sh1;               // origin of the pair's item range (fixed point) / shift of a 64-bin split of it
    int32_t hi1;                    // end of the pair's item range
    int32_t pad;
};

struct FastLayout {
    size_t slot_bytes;
    size_t off_hdr, off_rrot, off_aaf, off_bbf, off_aai, off_bbi, off_lo, off_w, off_cb, off_sh, off_cnt,
        off_cand, off_candd, off_rowpack, off_slist, off_scnt, off_samp_r, off_samp_c, off_live, off_nlive;
    int max_rows, max_cols, max_frames, lines, strips_c, slist_cap;
    int slog;                           // log2 of the diagonal sampling stride S
    int nst_r, nst_c;                   // sample slots per row (ceil(max_cols / S)) / per column
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

FastLayout make_layout(const SlotGeom &g, int max_frames) {
    FastLayout L;
    L.max_rows = g.max_rows; L.max_cols = g.max_cols; L.max_frames = max_frames;
    L.lines = g.max_rows + g.max_cols;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 128); return r; };
    L.off_hdr = take(sizeof(PairHdr));
    L.off_rrot = take((size_t)(max_frames + 8) * NBINS * 4);
    L.off_aaf = take((size_t)g.max_rows * 4);
    L.off_bbf = take((size_t)g.max_cols * 4);
    L.off_aai = take((size_t)g.max_rows * 4);
    L.off_bbi = take((size_t)g.max_cols * 4);
    L.off_lo = take((size_t)L.lines * 4);
    L.off_w = take((size_t)L.lines * 4);
    L.off_cb = take((size_t)L.lines * 4);
    L.off_sh = take((size_t)L.lines * 4);
    L.off_cnt = take((size_t)L.lines * 4);
    L.off_cand = take((size_t)L.lines * CAND_CAP * 2);
    L.off_candd = take((size_t)L.lines * CAND_CAP * 4);
    L.off_rowpack = take((size_t)g.max_rows * 16);
    L.strips_c = (g.max_cols + (32 * RCV - HALO) - 1) / (32 * RCV - HALO);
    L.slist_cap = 8192;                                   // uncertain cells one emit strip may record
    while (L.slist_cap < 4 * g.max_rows) L.slist_cap *= 2;
    L.off_slist = take((size_t)L.strips_c * L.slist_cap * 8);       // 2 words per uncertain cell: i | j << 14, item
    L.off_scnt = take((size_t)L.strips_c * 4);
    // diagonal sampling stride: the bracket a line gets from n_s samples holds ~2.35 L / sqrt(n_s) cells, which
    // the 64-bin split must bring under BRACKET_TARGET  =>  n_s >= (0.003 L)^2, S = L / n_s <= 1 / (9e-6 L)
    {
        const int Lmax = g.max_rows > g.max_cols ? g.max_rows : g.max_cols;
        const double smax = 1.0 / (9e-6 * (double)(Lmax > 1 ? Lmax : 1));
        L.slog = 5;
        while (L.slog > 2 && (double)(1 << L.slog) > smax) --L.slog;
        // short lines: keep at least 32 samples per line (a bracket from fewer samples is too wide to help;
        // measured at 500 frames: S = 16 or 8 give +30 % pairs/s over S = 32, at 2k frames S = 32 is best)
        while (L.slog > 2 && (Lmax >> L.slog) < 32) --L.slog;
        static const char *force = getenv("ACOSS_K2_SLOG");    // tuning switch
        if (force && force[0] >= '2' && force[0] <= '5') L.slog = force[0] - '0';
    }
    L.nst_r = ((g.max_cols - 1) >> L.slog) + 1;
    L.nst_c = ((g.max_rows - 1) >> L.slog) + 1;
    L.off_samp_r = take((size_t)L.nst_r * g.max_rows * 4);      // [t = j / S][i]
    L.off_samp_c = take((size_t)L.nst_c * g.max_cols * 4);      // [t = i / S][j]
    L.off_live = take((size_t)2 * SPARSE_CAP * 4);              // [side][SPARSE_CAP] lines still live after the dense level
    L.off_nlive = take(8);
    L.slot_bytes = align_up(o, 256);
    return L;
}

template <typename T>
__device__ __forceinline__ T *slot_ptr(char *base, const FastLayout &L, int slot, size_t off) {
    return reinterpret_cast<T *>(base + (size_t)slot * L.slot_bytes + off);
}

// ------------------------------------------------------------------------------------------------
// prep: rotated reference copy, exact float32 norms (reference order) + their fixed-point images,
// ranks, level-1 histogram range, zeroed candidate counters
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fast_prep_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                        const int32_t *__restrict__ oti, int64_t first,
                                                        FastLayout L, char *__restrict__ scratch, float qperc,
                                                        int guard, float fx_scale) {
    __shared__ int s_max[2];
    const int slot = blockIdx.x;
    const int64_t k = first + slot;
    const int q = pairs[2 * k], r = pairs[2 * k + 1], s = oti[k] % NBINS;
    const int nq = (int)(ts.offsets[q + 1] - ts.offsets[q]), nr = (int)(ts.offsets[r + 1] - ts.offsets[r]);
    const int Mx = nq - M9, Nx = nr - M9;
    const float *Q = ts.frames + ts.offsets[q] * NBINS;
    const float *R = ts.frames + ts.offsets[r] * NBINS;
    float *rrot = slot_ptr<float>(scratch, L, slot, L.off_rrot);
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0;
    for (int idx = threadIdx.x; idx < (nr + 8) * NBINS; idx += blockDim.x) {
        const int f = idx / NBINS, b = idx - f * NBINS;
        rrot[idx] = (f < nr) ? R[f * NBINS + rot_src(b, s)] : 0.f;
    }
    uint32_t *cnt = slot_ptr<uint32_t>(scratch, L, slot, L.off_cnt);
    for (int i = threadIdx.x; i < L.lines; i += blockDim.x) cnt[i] = 0u;
    for (int i = threadIdx.x; i < L.strips_c; i += blockDim.x) slot_ptr<uint32_t>(scratch, L, slot, L.off_scnt)[i] = 0u;
    if (threadIdx.x < 2) slot_ptr<uint32_t>(scratch, L, slot, L.off_nlive)[threadIdx.x] = 0u;
    __syncthreads();
    float *aaf = slot_ptr<float>(scratch, L, slot, L.off_aaf), *bbf = slot_ptr<float>(scratch, L, slot, L.off_bbf);
    int32_t *aai = slot_ptr<int32_t>(scratch, L, slot, L.off_aai), *bbi = slot_ptr<int32_t>(scratch, L, slot, L.off_bbi);
    int mxa = 0, mxb = 0;
    for (int i = threadIdx.x; i < Mx + Nx; i += blockDim.x) {
        const bool isq = i < Mx;
        const float *src = isq ? Q + (int64_t)i * NBINS : rrot + (int64_t)(i - Mx) * NBINS;
        double acc = 0.0;
        for (int t = 0; t < M9; ++t) {
            const float *fr = src + t * NBINS;
#pragma unroll
            for (int b = 0; b < NBINS; ++b) acc = acc_f32prod(acc, fr[b], fr[b]);
        }
        const float v = (float)acc;
        const int fx = __float2int_rn(v * fx_scale);
        if (isq) { aaf[i] = v; aai[i] = fx; mxa = max(mxa, fx); }
        else { bbf[i - Mx] = v; bbi[i - Mx] = fx; mxb = max(mxb, fx); }
    }
    atomicMax(&s_max[0], mxa);
    atomicMax(&s_max[1], mxb);
    __syncthreads();
    PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    int lo1 = -2 * EPS, sh1 = 0;
    const long long range1 = (long long)s_max[0] + s_max[1] + 4 * EPS + 1;
    while ((range1 >> sh1) > NBIN - 1) ++sh1;
    int fk2[2], ck2[2], quirk2[2];
    float kf2[2];
    for (int o = 0; o < 2; ++o) {
        const int Ln = (o == 0) ? Nx : Mx;                    // rows see Nx entries, columns see Mx
        const float kf = (Ln > 1) ? __fmul_rn((float)(Ln - 1), qperc) : __fmul_rn((float)Ln, qperc);
        const float fkf = floorf(kf), ckf = ceilf(kf);
        kf2[o] = kf; fk2[o] = (int)fkf; ck2[o] = min((int)ckf, Ln - 1);
        quirk2[o] = (fkf == ckf && !guard) ? 1 : 0;
    }
    if (threadIdx.x == 0) {
        h->nq = nq; h->nr = nr; h->Mx = Mx; h->Nx = Nx;
        for (int o = 0; o < 2; ++o) { h->fk[o] = fk2[o]; h->ck[o] = ck2[o]; h->quirk[o] = quirk2[o]; h->kf[o] = kf2[o]; }
        h->lo1 = lo1; h->sh1 = sh1; h->hi1 = lo1 + (int)range1;
    }
    // level-1 bracket state for every line; emit-ready defaults for quirk sides (threshold 0)
    int32_t *lo = slot_ptr<int32_t>(scratch, L, slot, L.off_lo), *w = slot_ptr<int32_t>(scratch, L, slot, L.off_w);
    int32_t *cb = slot_ptr<int32_t>(scratch, L, slot, L.off_cb), *sh = slot_ptr<int32_t>(scratch, L, slot, L.off_sh);
    int4 *rowpack = slot_ptr<int4>(scratch, L, slot, L.off_rowpack);
    for (int i = threadIdx.x; i < L.lines; i += blockDim.x) {
        const int side = (i < L.max_rows) ? 0 : 1;
        if (quirk2[side]) { lo[i] = 0; w[i] = 0; cb[i] = 0; sh[i] = -1; }
        else { lo[i] = lo1; w[i] = 0; cb[i] = 0; sh[i] = sh1; }
    }
    for (int i = threadIdx.x; i < Mx; i += blockDim.x) rowpack[i] = make_int4(aai[i], -2 * EPS, 4 * EPS, 0);
}

// ------------------------------------------------------------------------------------------------
// diagonal sampler: fixed-point items of the cells (i, j) with (j - i) % S == 0, i.e. every S-th
// diagonal.  A warp covers 24 consecutive positions of KD such diagonals: lane l computes the
// frame-level dot product of position i0 + l, the 9-tap window sum runs across lanes (4 shuffles).
// Every sample is a sample of its row and of its column: samp_r[j / S][i], samp_c[i / S][j].
// These samples only steer the histogram brackets (select kernel below); exactness never depends
// on them.
// ------------------------------------------------------------------------------------------------
constexpr int SKD = 4;              // diagonals per warp task
constexpr int SPOS = 32 - HALO;     // window sums one warp pass produces per diagonal (24)

__global__ void __launch_bounds__(128) fast_sample_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                          int64_t first, int n, FastLayout L,
                                                          char *__restrict__ scratch, int ngrp_max, int nblk_max,
                                                          float magic) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * 4 + warp;
    const int per_slot = ngrp_max * nblk_max;
    const int slot = (int)(task / per_slot);
    if (slot >= n) return;
    const int rem = (int)(task - (int64_t)slot * per_slot);
    const int grp = rem / nblk_max, blk = rem - grp * nblk_max;
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    if (h->quirk[0] && h->quirk[1]) return;
    const int nq = h->nq, nr = h->nr, Mx = h->Mx, Nx = h->Nx;
    const int slog = L.slog;
    const int i0 = blk * SPOS;
    if (i0 >= Mx) return;
    // diagonals d = S * u, u in [-(Mx-1)/S, (Nx-1)/S]
    const int u_lo = -((Mx - 1) >> slog);
    const int u0 = u_lo + grp * SKD;
    if (((int64_t)u0 << slog) > (int64_t)(Nx - 1) - i0) return;                       // block lies right of the matrix
    if ((((int64_t)u0 + SKD - 1) << slog) + i0 + SPOS - 1 < 0) return;                // ... or left of it
    const int64_t k = first + slot;
    const int q = pairs[2 * k];
    const float *Qf = ts.frames + ts.offsets[q] * NBINS;
    const float *Rf = slot_ptr<float>(scratch, L, slot, L.off_rrot);
    const int32_t *aai = slot_ptr<int32_t>(scratch, L, slot, L.off_aai), *bbi = slot_ptr<int32_t>(scratch, L, slot, L.off_bbi);
    int32_t *samp_r = slot_ptr<int32_t>(scratch, L, slot, L.off_samp_r), *samp_c = slot_ptr<int32_t>(scratch, L, slot, L.off_samp_c);
    const int i = i0 + lane;
    const bool iok = lane < SPOS && i < Mx;
    float x[NBINS];
    {
        const float4 *p = reinterpret_cast<const float4 *>(Qf + (int64_t)min(i, nq - 1) * NBINS);
        const float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
        x[0] = 2.f * v0.x; x[1] = 2.f * v0.y; x[2] = 2.f * v0.z; x[3] = 2.f * v0.w;
        x[4] = 2.f * v1.x; x[5] = 2.f * v1.y; x[6] = 2.f * v1.z; x[7] = 2.f * v1.w;
        x[8] = 2.f * v2.x; x[9] = 2.f * v2.y; x[10] = 2.f * v2.z; x[11] = 2.f * v2.w;
    }
    const int aa = iok ? aai[i] : 0;
    const int mbits = __float_as_int(magic);
#pragma unroll
    for (int kd = 0; kd < SKD; ++kd) {
        const int d = (u0 + kd) << slog;
        const int j = i + d;
        const float4 *p = reinterpret_cast<const float4 *>(Rf + (int64_t)min(max(j, 0), nr - 1) * NBINS);
        const float4 y0 = p[0], y1 = p[1], y2 = p[2];
        float acc = __fmaf_rn(x[0], y0.x, magic);
        acc = __fmaf_rn(x[1], y0.y, acc); acc = __fmaf_rn(x[2], y0.z, acc); acc = __fmaf_rn(x[3], y0.w, acc);
        acc = __fmaf_rn(x[4], y1.x, acc); acc = __fmaf_rn(x[5], y1.y, acc); acc = __fmaf_rn(x[6], y1.z, acc);
        acc = __fmaf_rn(x[7], y1.w, acc); acc = __fmaf_rn(x[8], y2.x, acc); acc = __fmaf_rn(x[9], y2.y, acc);
        acc = __fmaf_rn(x[10], y2.z, acc); acc = __fmaf_rn(x[11], y2.w, acc);
        const int v = __float_as_int(acc) - mbits;
        const int s1 = v + __shfl_down_sync(0xffffffffu, v, 1);
        const int s2 = s1 + __shfl_down_sync(0xffffffffu, s1, 2);
        const int s4 = s2 + __shfl_down_sync(0xffffffffu, s2, 4);
        const int T = s4 + __shfl_down_sync(0xffffffffu, v, 8);
        if (iok && j >= 0 && j < Nx) {
            const int z = aa + bbi[j] - T;
            samp_r[(int64_t)(j >> slog) * L.max_rows + i] = z;
            samp_c[(int64_t)(i >> slog) * L.max_cols + j] = z;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// select: one thread per line.  A private 256-bin histogram of the line's samples (over the pair's
// item range) gives the bins holding the sample ranks mu -+ 4.5 sigma of the wanted order statistic;
// that is the line's first bracket [lo, lo + 64 << sh) for the histogram sweeps.  A wrong bracket
// costs another sweep level, never a wrong result (under / overflow are counted exactly).
// ------------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 128;
constexpr int MIN_SAMPLES = 24;     // lines with fewer samples start from the whole item range

__global__ void __launch_bounds__(SEL_THREADS) fast_select_kernel(int n, FastLayout L, char *__restrict__ scratch) {
    extern __shared__ uint32_t s_sel_hist[];                 // [SBIN / 2][SEL_THREADS], two 16-bit counters per word
    const int slot = blockIdx.y;
    if (slot >= n) return;
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    const int Mx = h->Mx, Nx = h->Nx;
    const int gline = blockIdx.x * SEL_THREADS + threadIdx.x;
    if (blockIdx.x * SEL_THREADS >= Mx + Nx) return;
    uint32_t *hist = s_sel_hist + threadIdx.x;
#pragma unroll 4
    for (int b = 0; b < SBIN / 2; ++b) hist[b * SEL_THREADS] = 0u;
    if (gline >= Mx + Nx) return;
    const bool isrow = gline < Mx;
    const int idx = isrow ? gline : gline - Mx;
    const int side = isrow ? 0 : 1;
    if (h->quirk[side]) return;                               // threshold forced to 0: no selection on this side
    const int line = isrow ? idx : L.max_rows + idx;
    const int slog = L.slog, S = 1 << slog;
    const int Lother = isrow ? Nx : Mx;                       // entries of the line
    const int r0 = idx & (S - 1);                             // first sampled position along the line
    const int ns = (r0 < Lother) ? ((Lother - 1 - r0) >> slog) + 1 : 0;
    const int32_t *samp = isrow ? slot_ptr<int32_t>(scratch, L, slot, L.off_samp_r) : slot_ptr<int32_t>(scratch, L, slot, L.off_samp_c);
    const int64_t pitch = isrow ? L.max_rows : L.max_cols;
    const int lo1 = h->lo1, hi1 = h->hi1;
    int shs = 0;
    while ((((int64_t)hi1 - lo1) >> shs) > SBIN - 1) ++shs;
    for (int t0 = 0; t0 < ns; t0 += 8) {                      // 8 independent loads in flight per thread
        int z[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) z[u] = (t0 + u < ns) ? __ldg(samp + (int64_t)(t0 + u) * pitch + idx) : 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (t0 + u < ns) {
                const int b = min(max((z[u] - lo1) >> shs, 0), SBIN - 1);
                hist[(b >> 1) * SEL_THREADS] += 1u << (16 * (b & 1));
            }
        }
    }
    // wanted ranks floor(k) .. ceil(k) of Lother entries -> sample ranks mu -+ 4.5 sigma
    const float pq = ((float)h->fk[side] + 0.5f) / (float)Lother;
    const float mu = pq * (float)ns, sg = sqrtf(fmaxf(mu * (1.f - pq), 0.25f));
    const int r_lo = max((int)floorf(mu - 4.5f * sg) - 1, 0);
    const int r_hi = (int)ceilf(mu + 4.5f * sg) + 1;
    int cum = 0, b_lo = -1, b_hi = -1;
    for (int b = 0; b < SBIN / 2; ++b) {
        const uint32_t wv = hist[b * SEL_THREADS];
        const int c0 = wv & 0xffff, c1 = wv >> 16;
        if (b_lo < 0 && cum + c0 > r_lo) b_lo = 2 * b;
        if (b_hi < 0 && cum + c0 > r_hi) b_hi = 2 * b;
        cum += c0;
        if (b_lo < 0 && cum + c1 > r_lo) b_lo = 2 * b + 1;
        if (b_hi < 0 && cum + c1 > r_hi) b_hi = 2 * b + 1;
        cum += c1;
        if (b_hi >= 0) break;                                 // both ranks located (r_lo <= r_hi)
    }
    int lo = lo1, hi = hi1;
    // short lines have too few samples to steer anything: a 64-bin split of the whole item range already
    // isolates their order statistics (a bin then holds ~L / 40 cells)
    if (ns >= MIN_SAMPLES) {
        if (b_lo >= 0) lo = lo1 + (b_lo << shs);
        if (b_hi >= 0) hi = min(hi1, lo1 + ((b_hi + 1) << shs));
    }
    int sh = 0;
    while ((((int64_t)hi - lo) >> sh) > NBIN) ++sh;
    slot_ptr<int32_t>(scratch, L, slot, L.off_lo)[line] = lo;
    slot_ptr<int32_t>(scratch, L, slot, L.off_sh)[line] = sh;
}

// ------------------------------------------------------------------------------------------------
// the sweep: shared by the histogram kernels and the emit kernel
// ------------------------------------------------------------------------------------------------
template <int RC>
struct Sweep {
    static constexpr int COLS = 32 * RC;          // owned frames per strip
    static constexpr int OUTW = COLS - HALO;      // output windows per strip
    float y[RC][NBINS];                           // owned frames, pre-doubled
    int ring[M9][RC];                             // raw bits of the last 9 quantised e values
    int T[RC];                                    // fixed-point sliding diagonal sums (of 2e)
    unsigned okm[RC];                             // lane has a left neighbour holding column c-9
    unsigned nz0;                                 // all ones except on lane 0
    float magic;
    int mbits;

    __device__ __forceinline__ void init(const float *__restrict__ Y, int nY, int cb, int lane, float magic_) {
        magic = magic_; mbits = __float_as_int(magic_);
        nz0 = lane ? 0xffffffffu : 0u;
#pragma unroll
        for (int k = 0; k < RC; ++k) {
            const int c = cb + RC * lane + k;
            const float4 *p = reinterpret_cast<const float4 *>(Y + (int64_t)c * NBINS);
            float4 v0 = make_float4(0, 0, 0, 0), v1 = v0, v2 = v0;
            if (c < nY) { v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2); }
            y[k][0] = 2.f * v0.x; y[k][1] = 2.f * v0.y; y[k][2] = 2.f * v0.z; y[k][3] = 2.f * v0.w;
            y[k][4] = 2.f * v1.x; y[k][5] = 2.f * v1.y; y[k][6] = 2.f * v1.z; y[k][7] = 2.f * v1.w;
            y[k][8] = 2.f * v2.x; y[k][9] = 2.f * v2.y; y[k][10] = 2.f * v2.z; y[k][11] = 2.f * v2.w;
            T[k] = 0;
            const int kk = ((k - M9) % RC + RC) % RC;
            okm[k] = (lane >= (M9 - k + kk) / RC) ? 0xffffffffu : 0u;
#pragma unroll
            for (int u = 0; u < M9; ++u) ring[u][k] = mbits;
        }
    }

    // frame-level dot products of one streamed frame (12 floats in three float4) with the lane's RC owned
    // frames, quantised to fixed point (raw float bits of e + magic)
    __device__ __forceinline__ void dot(const float4 &x0, const float4 &x1, const float4 &x2, int (&eb)[RC]) const {
#pragma unroll
        for (int k = 0; k < RC; ++k) {
            // the chain starts at the magic constant, so every partial sum already sits in the fixed-point
            // binade (12 roundings of at most half a unit each: inside the EPS budget)
            float acc = __fmaf_rn(x0.x, y[k][0], magic);
            acc = __fmaf_rn(x0.y, y[k][1], acc); acc = __fmaf_rn(x0.z, y[k][2], acc); acc = __fmaf_rn(x0.w, y[k][3], acc);
            acc = __fmaf_rn(x1.x, y[k][4], acc); acc = __fmaf_rn(x1.y, y[k][5], acc); acc = __fmaf_rn(x1.z, y[k][6], acc);
            acc = __fmaf_rn(x1.w, y[k][7], acc); acc = __fmaf_rn(x2.x, y[k][8], acc); acc = __fmaf_rn(x2.y, y[k][9], acc);
            acc = __fmaf_rn(x2.z, y[k][10], acc); acc = __fmaf_rn(x2.w, y[k][11], acc);
            eb[k] = __float_as_int(acc);
        }
    }

    // slides the diagonal sums T by one row.  U = a % 9 (static ring slot).
    template <int U>
    __device__ __forceinline__ void finish(const int (&eb)[RC]) {
        int old[RC];
        // e[a-9][c-9]: column c-9 lives dl lanes to the left, in register (k - 9) mod RC of ring slot U
#pragma unroll
        for (int k = 0; k < RC; ++k) {
            const int kk = ((k - M9) % RC + RC) % RC;
            const int dl = (M9 - k + kk) / RC;
            const unsigned v = (unsigned)__shfl_up_sync(0xffffffffu, ring[U][kk], dl);
            old[k] = (int)((v & okm[k]) | ((unsigned)mbits & ~okm[k]));
        }
        const int tl = (int)((unsigned)__shfl_up_sync(0xffffffffu, T[RC - 1], 1) & nz0);
#pragma unroll
        for (int k = RC - 1; k >= 1; --k) T[k] = T[k - 1] + eb[k] - old[k];
        T[0] = tl + eb[0] - old[0];
#pragma unroll
        for (int k = 0; k < RC; ++k) ring[U][k] = eb[k];
    }
};

template <int V> struct IC { static constexpr int value = V; };
template <bool V> struct BC { static constexpr bool value = V; };

// Drives a sweep over streamed frames 0 .. nrows-1 (nrows >= 10 always: Mx >= 2).  The next frame is
// loaded into the frame registers as soon as the dot products have consumed them, the next row's
// parameter word right after its use: loads overlap the rest of the row, no register copies.  fn(a, param) runs for
// rows a >= HALO, param = P[a - HALO].
template <int RC, typename PT, typename Fn>
__device__ __forceinline__ void run_sweep(Sweep<RC> &sw, const float *__restrict__ X, const PT *__restrict__ P,
                                          int nrows, Fn &&fn) {
    const float4 *px = reinterpret_cast<const float4 *>(X);
    float4 c0 = __ldg(px), c1 = __ldg(px + 1), c2 = __ldg(px + 2);
    PT pc{};
    const PT *pp = P - HALO;                       // pp[a] is the parameter of row a
    int a = 0;
    auto step = [&](auto uc, auto callc, auto pfc) {
        constexpr int U = decltype(uc)::value;
        constexpr bool CALL = decltype(callc)::value, PF = decltype(pfc)::value;
        int eb[RC];
        PT pn{};
        if (PF) pn = __ldg(pp + a + 1);            // next row's parameter: a whole row ahead of its use
        sw.dot(c0, c1, c2, eb);
        // the frame registers are dead now: refill them with the next row while this row finishes
        px += 3;
        c0 = __ldg(px); c1 = __ldg(px + 1); c2 = __ldg(px + 2);
        sw.template finish<U>(eb);
        if (CALL) fn(a, pc);
        if (PF) pc = pn;
        ++a;
    };
    // rows 0..8: the diagonal sums fill up; only row 8 produces a window
    step(IC<0>{}, BC<false>{}, BC<false>{}); step(IC<1>{}, BC<false>{}, BC<false>{}); step(IC<2>{}, BC<false>{}, BC<false>{});
    step(IC<3>{}, BC<false>{}, BC<false>{}); step(IC<4>{}, BC<false>{}, BC<false>{}); step(IC<5>{}, BC<false>{}, BC<false>{});
    step(IC<6>{}, BC<false>{}, BC<false>{}); step(IC<7>{}, BC<false>{}, BC<true>{}); step(IC<8>{}, BC<true>{}, BC<true>{});
#pragma unroll 1
    while (a + M9 <= nrows) {
        step(IC<0>{}, BC<true>{}, BC<true>{}); step(IC<1>{}, BC<true>{}, BC<true>{}); step(IC<2>{}, BC<true>{}, BC<true>{});
        step(IC<3>{}, BC<true>{}, BC<true>{}); step(IC<4>{}, BC<true>{}, BC<true>{}); step(IC<5>{}, BC<true>{}, BC<true>{});
        step(IC<6>{}, BC<true>{}, BC<true>{}); step(IC<7>{}, BC<true>{}, BC<true>{}); step(IC<8>{}, BC<true>{}, BC<true>{});
    }
    const int rem = nrows - a;                     // 0..8 rows left, ring slots 0..rem-1
    if (rem > 0) step(IC<0>{}, BC<true>{}, BC<true>{});
    if (rem > 1) step(IC<1>{}, BC<true>{}, BC<true>{});
    if (rem > 2) step(IC<2>{}, BC<true>{}, BC<true>{});
    if (rem > 3) step(IC<3>{}, BC<true>{}, BC<true>{});
    if (rem > 4) step(IC<4>{}, BC<true>{}, BC<true>{});
    if (rem > 5) step(IC<5>{}, BC<true>{}, BC<true>{});
    if (rem > 6) step(IC<6>{}, BC<true>{}, BC<true>{});
    if (rem > 7) step(IC<7>{}, BC<true>{}, BC<true>{});
}

// ------------------------------------------------------------------------------------------------
// Bracket update shared by the dense and the sparse histogram levels.  hist[b * stride] (shifted / masked)
// is the count of bin b: b = 0 items below the bracket, 1..NBIN its bins, NBIN + 1 items above.  The new
// bracket is the run of bins holding ranks fk and ck (the under / overflow "bins" reach to the end of the
// pair's item range, the next level splits them).
// ------------------------------------------------------------------------------------------------
// cells a finished bracket may hold: BRACKET_TARGET on long lines, a sixteenth of the line on short ones (every
// bracket cell is an uncertain cell of the emit sweep, and a short line has few cells to begin with)
__device__ __forceinline__ int bracket_target(int line_len) { return min(BRACKET_TARGET, max(6, line_len >> 4)); }

struct Bracket {
    int lo, w, below, sh;       // new origin, width, exact count of items below lo, shift of the next level
    bool done, miss, bad;
};

template <int STRIDE>
__device__ __forceinline__ Bracket split_bracket(const uint32_t *hp, int hs, uint32_t mask, int fk, int ck, int lo, int sh,
                                                 int rlo, int rhi, int target) {
    int cum = 0, b1 = -2, b2 = -2, cb1 = 0, cend = 0;          // the underflow bin counts every item below the bracket
    for (int b = 0; b < NBIN + 2; ++b) {                      // b - 1 = bin of the bracket; -1 under, NBIN over
        const int c = (int)((hp[b * STRIDE] >> hs) & mask);
        if (b1 < -1 && cum + c > fk) { b1 = b - 1; cb1 = cum; }
        if (b2 < -1 && cum + c > ck) { b2 = b - 1; cend = cum + c; }
        cum += c;
    }
    Bracket r;
    r.bad = (b1 < -1 || b2 < -1);
    const long long nlo = (b1 < 0) ? (long long)rlo : (long long)lo + ((long long)b1 << sh);
    const long long nhi = (b2 >= NBIN) ? (long long)rhi : (b2 < 0) ? (long long)lo : (long long)lo + ((long long)(b2 + 1) << sh);
    r.miss = (b1 < 0) || (b2 >= NBIN);
    const long long range = (nhi > nlo) ? nhi - nlo : 1;
    int sh2 = 0;
    while ((range >> sh2) > NBIN) ++sh2;
    // split again unless the bracket is small enough or cannot shrink
    r.done = !r.miss && ((cend - cb1 <= target) || (sh == 0));
    r.lo = (int)nlo; r.w = (int)range; r.below = cb1; r.sh = sh2;
    return r;
}

// ------------------------------------------------------------------------------------------------
// histogram sweep.  ORIENT = 0: owned = reference columns (column thresholds), streamed = query.
//                   ORIENT = 1: owned = query rows (row thresholds), streamed = rotated reference.
// Every live line enters with a bracket [lo, lo + 64 << sh) (first level: from the sample selection).  The
// sweep counts the line's items into 64 bins of the bracket plus an underflow and an overflow bin, so the
// position of the wanted ranks is known exactly whatever the bracket was: inside (bracket shrinks to the
// bins holding floor(k) .. ceil(k)), below or above (bracket moves to that side of the pair's item range
// and the next level splits it).
// ------------------------------------------------------------------------------------------------
template <int RC, int ORIENT>
__global__ void __launch_bounds__(32 * WPC, 3) fast_hist_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                             int64_t first, int n, FastLayout L,
                                                             char *__restrict__ scratch, int strips_max, float magic,
                                                             uint32_t *__restrict__ status, uint32_t *__restrict__ dbg,
                                                             int min_live, int final_level) {
    extern __shared__ uint32_t s_hist[];                      // [WPC][NBIN + 2][RC / 2][32], two 16-bit counters per word
    using SW = Sweep<RC>;
    static_assert(RC % 2 == 0, "histogram packing needs an even number of register columns");
    constexpr int HW = RC / 2;                                // words per (bin, lane)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * WPC + warp;
    const int slot = (int)(task / strips_max), strip = (int)(task % strips_max);
    if (slot >= n) return;
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    if (h->quirk[ORIENT == 0 ? 1 : 0]) return;               // threshold forced to 0: nothing to select
    const int64_t k = first + slot;
    const int q = pairs[2 * k];
    const int nY = (ORIENT == 0) ? h->nr : h->nq, nX = (ORIENT == 0) ? h->nq : h->nr;
    const int My = nY - M9;                                   // owned windows
    const int cb = strip * SW::OUTW;
    if (cb >= My) return;
    const float *Qf = ts.frames + ts.offsets[q] * NBINS;
    const float *Rf = slot_ptr<float>(scratch, L, slot, L.off_rrot);
    const float *Y = (ORIENT == 0) ? Rf : Qf, *X = (ORIENT == 0) ? Qf : Rf;
    const int32_t *xn = slot_ptr<int32_t>(scratch, L, slot, ORIENT == 0 ? L.off_aai : L.off_bbi);
    const int32_t *yn = slot_ptr<int32_t>(scratch, L, slot, ORIENT == 0 ? L.off_bbi : L.off_aai);
    const int line0 = (ORIENT == 0) ? L.max_rows : 0;         // first line index of the owned side
    int32_t *lo_a = slot_ptr<int32_t>(scratch, L, slot, L.off_lo) + line0;
    int32_t *w_a = slot_ptr<int32_t>(scratch, L, slot, L.off_w) + line0;
    int32_t *cb_a = slot_ptr<int32_t>(scratch, L, slot, L.off_cb) + line0;
    int32_t *sh_a = slot_ptr<int32_t>(scratch, L, slot, L.off_sh) + line0;

    // a line is live while its bracket still holds more than BRACKET_TARGET cells and can be split further
    // (shift >= 0 in sh_a; -1 = done).  Strips with nothing left to do return at once, so the extra levels
    // cost nothing on ordinary pairs.
    int ynrel[RC], shf[RC];
    bool valid[RC];
    bool work = false;
#pragma unroll
    for (int kk = 0; kk < RC; ++kk) {
        const int j = cb + RC * lane + kk - HALO;            // owned window of this register column
        valid[kk] = (j >= cb) && (j < My) && (j < cb + SW::OUTW);
        const int shv = valid[kk] ? sh_a[j] : -1;
        if (shv < 0) valid[kk] = false;
        shf[kk] = valid[kk] ? shv : 0;
        // bin index = ((z - lo) >> sh) + 1 clamped to [0, NBIN + 1]: 0 = below the bracket, NBIN + 1 = above
        ynrel[kk] = valid[kk] ? yn[j] - lo_a[j] + (1 << shv) : 0x40000000;   // idle columns land in the overflow bin
        work |= valid[kk];
    }
    if (!__any_sync(0xffffffffu, work)) return;
    int n_live = 0, n_miss = 0, n_left = 0;                  // diagnostics (acoss_debug_counters)
#pragma unroll
    for (int kk = 0; kk < RC; ++kk) n_live += valid[kk] ? 1 : 0;
    const int side = (ORIENT == 0) ? 1 : 0;
    uint32_t *nlive = slot_ptr<uint32_t>(scratch, L, slot, L.off_nlive) + side;
    int32_t *live = slot_ptr<int32_t>(scratch, L, slot, L.off_live) + side * SPARSE_CAP;
    if (__reduce_add_sync(0xffffffffu, n_live) < min_live) {
        // too few live lines to pay for a dense sweep of the strip: they go to the sparse refinement as they are
#pragma unroll
        for (int kk = 0; kk < RC; ++kk) {
            if (!valid[kk]) continue;
            const unsigned pos = atomicAdd(nlive, 1u);
            if (pos < SPARSE_CAP) live[pos] = cb + RC * lane + kk - HALO;
        }
        return;
    }
    SW sw;
    sw.init(Y, nY, cb, lane, magic);
    uint32_t *hist = s_hist + (size_t)warp * (NBIN + 2) * HW * 32 + lane;
#pragma unroll 1
    for (int b = 0; b < (NBIN + 2) * HW; ++b) hist[b * 32] = 0u;
    __syncwarp();

    const int nrows = nX - 1;                                 // streamed frames 0 .. nX-2 (F4: last frame unused)
    run_sweep<RC, int>(sw, X, xn, nrows, [&](int, int xb) {
#pragma unroll
        for (int kk = 0; kk < RC; ++kk) {
            const int zr = xb + ynrel[kk] - sw.T[kk];
            const int idx = __vimin_s32_relu(zr >> shf[kk], NBIN + 1);
            atomicAdd(&hist[(idx * HW + kk / 2) * 32], 1u << (16 * (kk & 1)));   // thread-private bank: conflict-free
        }
    });
    __syncwarp();
    // per-thread scan of its own columns' histograms
    const int fk = h->fk[side], ck = h->ck[side];
    const int rlo = h->lo1, rhi = h->hi1;
#pragma unroll
    for (int kk = 0; kk < RC; ++kk) {
        if (!valid[kk]) continue;
        const int j = cb + RC * lane + kk - HALO;
        const Bracket br = split_bracket<HW * 32>(hist + (kk / 2) * 32, 16 * (kk & 1), 0xffffu, fk, ck, lo_a[j], shf[kk], rlo, rhi,
                                                  bracket_target(nX - M9));
        if (br.bad) {                                         // cannot happen: the bins cover every item of the line
            atomicOr(&status[k], PAIR_ST_FALLBACK | 4u);      // reason 4: rank not found
            sh_a[j] = -1;
            continue;
        }
        lo_a[j] = br.lo;
        w_a[j] = br.w;
        cb_a[j] = br.below;
        sh_a[j] = br.done ? -1 : br.sh;
        n_miss += br.miss ? 1 : 0;
        n_left += br.done ? 0 : 1;
        if (ORIENT == 1) {
            int4 *rowpack = slot_ptr<int4>(scratch, L, slot, L.off_rowpack);
            rowpack[j] = make_int4(yn[j], br.lo - 2 * EPS, br.w + 4 * EPS, 0);
        }
        if (!br.done && final_level) {                        // the sparse level refines it
            const unsigned pos = atomicAdd(nlive, 1u);
            if (pos < SPARSE_CAP) live[pos] = j;
        }
    }
    n_live = __reduce_add_sync(0xffffffffu, n_live);
    n_miss = __reduce_add_sync(0xffffffffu, n_miss);
    n_left = __reduce_add_sync(0xffffffffu, n_left);
    if (lane == 0) {
        atomicAdd(&dbg[0], 1u); atomicAdd(&dbg[1], (unsigned)n_live);
        if (n_miss) atomicAdd(&dbg[2], (unsigned)n_miss);
        if (n_left) atomicAdd(&dbg[3], (unsigned)n_left);
    }
}

// ------------------------------------------------------------------------------------------------
// sparse refinement: the few lines whose bracket is still crowded after the dense level (or whose ranks
// fell outside the sampled bracket).  One lane owns one line: the 9 owned frames of its window stay in
// registers, every streamed frame feeds the 9 rows in flight (acc[(a - t) % 9] += e(a, t)), one row
// completes per step.  The quantised dot products are the same FMA chains as in the sweeps, so the items
// are identical.  A warp repeats the sweep until all its lines are done (at most SPARSE_LEVELS times).
// ------------------------------------------------------------------------------------------------
constexpr int SPARSE_LEVELS = 3;
constexpr int DENSE2_MIN_LIVE = 16;  // live lines a strip must hold for the second dense level to sweep it

template <int U>
__device__ __forceinline__ void sparse_step(const float (&y)[M9][NBINS], int (&acc)[M9], const float4 &x0, const float4 &x1,
                                            const float4 &x2, float magic) {
#pragma unroll
    for (int t = 0; t < M9; ++t) {
        const float *yt = y[t];
        float a = __fmaf_rn(x0.x, yt[0], magic);
        a = __fmaf_rn(x0.y, yt[1], a); a = __fmaf_rn(x0.z, yt[2], a); a = __fmaf_rn(x0.w, yt[3], a);
        a = __fmaf_rn(x1.x, yt[4], a); a = __fmaf_rn(x1.y, yt[5], a); a = __fmaf_rn(x1.z, yt[6], a);
        a = __fmaf_rn(x1.w, yt[7], a); a = __fmaf_rn(x2.x, yt[8], a); a = __fmaf_rn(x2.y, yt[9], a);
        a = __fmaf_rn(x2.z, yt[10], a); a = __fmaf_rn(x2.w, yt[11], a);
        constexpr int dummy = 0; (void)dummy;
        const int slot = ((U - t) % M9 + M9) % M9;            // row a - t lives in slot (a - t) % 9
        if (t == 0) acc[slot] = __float_as_int(a);
        else acc[slot] += __float_as_int(a);
    }
}

__global__ void __launch_bounds__(32 * WPC, 2) fast_sparse_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                                  int64_t first, int n, FastLayout L,
                                                                  char *__restrict__ scratch, float magic,
                                                                  uint32_t *__restrict__ status, uint32_t *__restrict__ dbg) {
    __shared__ uint32_t s_sp[WPC][NBIN + 2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int CHUNKS = SPARSE_CAP / 32;
    const int64_t task = (int64_t)blockIdx.x * WPC + warp;
    const int slot = (int)(task / (2 * CHUNKS));
    if (slot >= n) return;
    const int rem = (int)(task - (int64_t)slot * (2 * CHUNKS));
    const int side = rem / CHUNKS, chunk = rem - side * CHUNKS;    // side 0: rows (owned = query), 1: columns
    const uint32_t cnt_all = slot_ptr<uint32_t>(scratch, L, slot, L.off_nlive)[side];
    const int64_t k = first + slot;
    if (cnt_all > SPARSE_CAP && chunk == 0 && lane == 0) atomicOr(&status[k], PAIR_ST_FALLBACK | 8u);   // reason 8: too many crowded lines
    const int cnt = (int)min(cnt_all, (uint32_t)SPARSE_CAP);
    if (chunk * 32 >= cnt) return;
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    const int q = pairs[2 * k];
    const int nY = side ? h->nr : h->nq, nX = side ? h->nq : h->nr;
    const float *Qf = ts.frames + ts.offsets[q] * NBINS;
    const float *Rf = slot_ptr<float>(scratch, L, slot, L.off_rrot);
    const float *Y = side ? Rf : Qf, *X = side ? Qf : Rf;
    const int32_t *xn = slot_ptr<int32_t>(scratch, L, slot, side ? L.off_aai : L.off_bbi);
    const int32_t *yn = slot_ptr<int32_t>(scratch, L, slot, side ? L.off_bbi : L.off_aai);
    const int line0 = side ? L.max_rows : 0;
    int32_t *lo_a = slot_ptr<int32_t>(scratch, L, slot, L.off_lo) + line0;
    int32_t *w_a = slot_ptr<int32_t>(scratch, L, slot, L.off_w) + line0;
    int32_t *cb_a = slot_ptr<int32_t>(scratch, L, slot, L.off_cb) + line0;
    int32_t *sh_a = slot_ptr<int32_t>(scratch, L, slot, L.off_sh) + line0;
    const int li = chunk * 32 + lane;
    bool livel = li < cnt;
    const int j = livel ? slot_ptr<int32_t>(scratch, L, slot, L.off_live)[side * SPARSE_CAP + li] : 0;
    // owned window j: frames j .. j+8, pre-doubled (the sweeps double the owned side too)
    float y[M9][NBINS];
#pragma unroll
    for (int t = 0; t < M9; ++t) {
        const float4 *p = reinterpret_cast<const float4 *>(Y + (int64_t)min(j + t, nY - 1) * NBINS);
        const float4 v0 = p[0], v1 = p[1], v2 = p[2];
        y[t][0] = 2.f * v0.x; y[t][1] = 2.f * v0.y; y[t][2] = 2.f * v0.z; y[t][3] = 2.f * v0.w;
        y[t][4] = 2.f * v1.x; y[t][5] = 2.f * v1.y; y[t][6] = 2.f * v1.z; y[t][7] = 2.f * v1.w;
        y[t][8] = 2.f * v2.x; y[t][9] = 2.f * v2.y; y[t][10] = 2.f * v2.z; y[t][11] = 2.f * v2.w;
    }
    const int ynj = livel ? yn[j] : 0;
    const int fk = h->fk[side], ck = h->ck[side], rlo = h->lo1, rhi = h->hi1;
    const int mb9 = M9 * __float_as_int(magic);               // the 9 magic offsets inside a completed sum
    uint32_t *hist = &s_sp[warp][0][lane];
    int lo = livel ? lo_a[j] : 0, sh = livel ? sh_a[j] : 0;
    if (livel && sh < 0) livel = false;
    const int nrows = nX - 1;                                 // streamed frames 0 .. nX-2
    int n_swept = 0;
    for (int lvl = 0; lvl < SPARSE_LEVELS && __any_sync(0xffffffffu, livel); ++lvl) {
        ++n_swept;
#pragma unroll 1
        for (int b = 0; b < NBIN + 2; ++b) hist[b * 32] = 0u;
        const int yrel = livel ? ynj - lo + (1 << sh) : 0x40000000;   // idle lanes land in the overflow bin
        const float4 *px = reinterpret_cast<const float4 *>(X);
        float4 c0 = __ldg(px), c1 = __ldg(px + 1), c2 = __ldg(px + 2);
        int acc[M9];
#pragma unroll
        for (int u = 0; u < M9; ++u) acc[u] = 0;
        int a = 0;
        auto step = [&](auto uc) {
            constexpr int U = decltype(uc)::value;
            sparse_step<U>(y, acc, c0, c1, c2, magic);
            px += 3;
            c0 = __ldg(px); c1 = __ldg(px + 1); c2 = __ldg(px + 2);
            if (a >= HALO) {                                  // row a - 8 is complete (slot (U + 1) % 9)
                const int T = acc[(U + 1) % M9] - mb9;
                const int zr = __ldg(xn + a - HALO) + yrel - T;
                const int idx = __vimin_s32_relu(zr >> sh, NBIN + 1);
                atomicAdd(&hist[idx * 32], 1u);
            }
            ++a;
        };
#pragma unroll 1
        while (a + M9 <= nrows) {
            step(IC<0>{}); step(IC<1>{}); step(IC<2>{}); step(IC<3>{}); step(IC<4>{});
            step(IC<5>{}); step(IC<6>{}); step(IC<7>{}); step(IC<8>{});
        }
        const int remr = nrows - a;
        if (remr > 0) step(IC<0>{});
        if (remr > 1) step(IC<1>{});
        if (remr > 2) step(IC<2>{});
        if (remr > 3) step(IC<3>{});
        if (remr > 4) step(IC<4>{});
        if (remr > 5) step(IC<5>{});
        if (remr > 6) step(IC<6>{});
        if (remr > 7) step(IC<7>{});
        __syncwarp();
        if (livel) {
            const Bracket br = split_bracket<32>(hist, 0, 0xffffffffu, fk, ck, lo, sh, rlo, rhi, bracket_target(nX - M9));
            if (br.bad) { atomicOr(&status[k], PAIR_ST_FALLBACK | 4u); livel = false; sh_a[j] = -1; }
            else {
                lo = br.lo; sh = br.sh;
                lo_a[j] = br.lo; w_a[j] = br.w; cb_a[j] = br.below; sh_a[j] = br.done ? -1 : br.sh;
                if (side == 0) slot_ptr<int4>(scratch, L, slot, L.off_rowpack)[j] = make_int4(ynj, br.lo - 2 * EPS, br.w + 4 * EPS, 0);
                if (br.done) livel = false;
            }
        }
        __syncwarp();
    }
    if (lane == 0) { atomicAdd(&dbg[0], 1u); atomicAdd(&dbg[1], (unsigned)n_swept); }
    const unsigned left = __ballot_sync(0xffffffffu, livel);
    if (lane == 0 && left) atomicAdd(&dbg[2], (unsigned)__popc(left));
}

// ------------------------------------------------------------------------------------------------
// emit sweep (orientation 0: owned = reference columns, streamed = query rows)
// ------------------------------------------------------------------------------------------------
template <int RC>
__global__ void __launch_bounds__(32 * WPC, 3) fast_emit_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                             int64_t first, int n, FastLayout L,
                                                             char *__restrict__ scratch, int strips_max, float magic,
                                                             uint32_t *__restrict__ crp_all, int words,
                                                             int64_t crp_words) {
    using SW = Sweep<RC>;
    static_assert(RC == 4, "emit word assembly: 4 bits per lane that never straddle a word, groups of <= 8 lanes");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * WPC + warp;
    const int slot = (int)(task / strips_max), strip = (int)(task % strips_max);
    if (slot >= n) return;
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    const int64_t k = first + slot;
    const int q = pairs[2 * k];
    const int nY = h->nr, nX = h->nq, My = nY - M9, Mx = nX - M9;
    const int cb = strip * SW::OUTW;
    if (cb >= My) return;
    const float *X = ts.frames + ts.offsets[q] * NBINS;
    const float *Y = slot_ptr<float>(scratch, L, slot, L.off_rrot);
    const int32_t *yn = slot_ptr<int32_t>(scratch, L, slot, L.off_bbi);
    const int4 *rowpack = slot_ptr<int4>(scratch, L, slot, L.off_rowpack);
    const int32_t *lo_c = slot_ptr<int32_t>(scratch, L, slot, L.off_lo) + L.max_rows;
    const int32_t *w_c = slot_ptr<int32_t>(scratch, L, slot, L.off_w) + L.max_rows;
    // list of uncertain cells of this strip: 2 words per cell (i | j << 14, fixed-point item).  Every lane fills
    // chunks of LCH entries it takes from the strip's pool (one returning atomic per chunk), so an append is a
    // plain lane-private store: no warp cooperation in the sweep
    constexpr unsigned LCH = 16;
    const unsigned lcap = (unsigned)L.slist_cap;
    uint2 *pool = reinterpret_cast<uint2 *>(slot_ptr<uint32_t>(scratch, L, slot, L.off_slist)) + (size_t)strip * lcap;
    uint32_t *pool_ctr = slot_ptr<uint32_t>(scratch, L, slot, L.off_scnt) + strip;
    unsigned lp = 0u, lend = 0u;
    uint32_t *crp = crp_all + (int64_t)slot * crp_words;

    SW sw;
    sw.init(Y, nY, cb, lane, magic);
    int ynv[RC], ycl[RC];                                     // bb_fix[j], bb_fix[j] - (colLo - 2 EPS)
    unsigned cw1[RC];                                         // colW + 4 EPS - 1
#pragma unroll
    for (int kk = 0; kk < RC; ++kk) {
        const int j = cb + RC * lane + kk - HALO;
        const bool valid = (j >= cb) && (j < My) && (j < cb + SW::OUTW);
        ynv[kk] = valid ? yn[j] : 0x20000000;                 // invalid: item huge => never in, never uncertain
        ycl[kk] = valid ? yn[j] - (lo_c[j] - 2 * EPS) : 0x20000000;
        cw1[kk] = valid ? (unsigned)(w_c[j] + 4 * EPS - 1) : 0u;
    }
    // output placement: strip bit t = RC*lane + kk - HALO <-> CRP column cb + t.  cb is a multiple of 8 and the
    // lane's RC bits start at a multiple of RC, so they never straddle a 32-bit word: every lane contributes
    // a nibble to word lword of the row at a position that is constant over the sweep.
    const unsigned ij0 = (unsigned)(cb + RC * lane - HALO) << 14;   // column part of a list entry (+ kk << 14)
    const int gpos = (cb & 31) + RC * lane - HALO;
    const int lword = gpos >> 5;                              // -1 for halo lanes (their columns are invalid: nibble 0)
    const int lbit = gpos & 31;
    const unsigned full = 0xffffffffu;
    // lanes that feed the same CRP word form a group of consecutive lanes; a 3-step segmented OR (shuffle down,
    // masked by group membership) leaves the complete word in the lowest lane of each group, which owns the store
    const unsigned gmask = __match_any_sync(full, lword);
    const bool gleader = (lword >= 0) && ((gmask & ((1u << lane) - 1u)) == 0u);
    const unsigned m1 = (lane + 1 < 32 && ((gmask >> (lane + 1)) & 1u)) ? 0xffffffffu : 0u;
    const unsigned m2 = (lane + 2 < 32 && ((gmask >> (lane + 2)) & 1u)) ? 0xffffffffu : 0u;
    const unsigned m4 = (lane + 4 < 32 && ((gmask >> (lane + 4)) & 1u)) ? 0xffffffffu : 0u;
    // advances by `words` per row; two rows behind the sweep (segmented-OR pipeline).  Slot 0's rows -2, -1 would
    // lie before the CRP buffer, but nothing is stored for them (their words are zero)
    uint32_t *rowp = crp + (cb >> 5) + (lword >= 0 ? lword : 0) - 2 * (int64_t)words;
    unsigned p1 = 0u, p2 = 0u;
    const int nrows = nX - 1;
    // Per cell, with z the fixed-point item (sign-bit / unsigned-compare arithmetic):
    //   ar = z - (rowLo - 2 EPS), ac = z - (colLo - 2 EPS)
    //   certainly in  <=> ar < 0 and ac < 0                                  sign(ar & ac)
    //   row zone      <=> 0 <= ar <= rw1 (rw1 = rowW + 4 EPS - 1)            (unsigned)ar <= rw1
    //   near zero     <=> z < 2 EPS  <=>  ar < 4 EPS - rowLo                 (always evaluated exactly: F7's NaN)
    //   uncertain     <=> row zone or column zone or near zero
    // A near-zero cell that is certainly in is emitted as 1 and listed as well: its exact evaluation either
    // confirms a tiny distance (<= both thresholds, resolve checks thr against the zone) or raises the NaN error.
    run_sweep<RC, int4>(sw, X, rowpack, nrows, [&](int a, const int4 &rp) {   // rp = {aa_fix, rowLo - 2EPS, rowW + 4EPS, -}
        const int xr = rp.x - rp.y, nz = 2 * EPS - rp.y;
        const unsigned rw1 = (unsigned)(rp.z - 1);
        int ar[RC];
        bool unc[RC];
        unsigned nib = 0u;
#pragma unroll
        for (int kk = RC - 1; kk >= 0; --kk) {
            ar[kk] = xr + ynv[kk] - sw.T[kk];
            const int ac = rp.x + ycl[kk] - sw.T[kk];
            unc[kk] = ((unsigned)ar[kk] <= rw1) || ((unsigned)ac <= cw1[kk]) || (ar[kk] < nz);
            nib = __funnelshift_l((unsigned)(ar[kk] & ac), nib, 1);   // sign bit -> bit 0, earlier cells move up
        }
        // the three steps of the segmented OR run on three consecutive rows (independent shuffles, no chain):
        // this row enters step 1, row a-1 step 2, row a-2 step 4 and is stored
        const unsigned v0 = nib << lbit;
        const unsigned t1 = __shfl_down_sync(full, v0, 1), t2 = __shfl_down_sync(full, p1, 2),
                       t4 = __shfl_down_sync(full, p2, 4);
        const unsigned vout = p2 | (t4 & m4);
        if (gleader && vout) atomicOr(rowp, vout);            // rows -2, -1 of the pipeline are all zero
        p2 = p1 | (t2 & m2);
        p1 = v0 | (t1 & m1);
        rowp += words;
        if (__any_sync(full, (unc[0] || unc[1]) || (unc[2] || unc[3]))) {
            // uncertain cells go to the lane's own list (no warp cooperation); fast_scatter_kernel classifies them
            const unsigned i = (unsigned)(a - HALO);          // query window (CRP row); i < Mx by construction
#pragma unroll
            for (int kk = 0; kk < RC; ++kk) {
                if (unc[kk]) {
                    if (lp == lend) { lp = atomicAdd(pool_ctr, LCH); lend = lp + LCH; }
                    if (lp < lcap) pool[lp] = make_uint2(i | (ij0 + ((unsigned)kk << 14)), (unsigned)(ar[kk] + rp.y));
                    ++lp;
                }
            }
        }
    });
    (void)Mx;
    {   // drain the segmented-OR pipeline: rows Mx-2 and Mx-1
        const unsigned t2 = __shfl_down_sync(full, p1, 2), t4 = __shfl_down_sync(full, p2, 4);
        const unsigned vout = p2 | (t4 & m4);
        if (gleader && vout) atomicOr(rowp, vout);
        p2 = p1 | (t2 & m2);
        rowp += words;
        const unsigned t4b = __shfl_down_sync(full, p2, 4);
        const unsigned vout2 = p2 | (t4b & m4);
        if (gleader && vout2) atomicOr(rowp, vout2);
    }
    for (; lp < lend; ++lp)                                   // unused tail of the lane's last chunk
        if (lp < lcap) pool[lp] = make_uint2(0xffffffffu, 0u);
}

// ------------------------------------------------------------------------------------------------
// scatter: lane lists -> per-row / per-column candidate lists (the returning atomics live here, in a
// kernel with enough parallelism to hide them)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fast_scatter_kernel(int n, FastLayout L, char *__restrict__ scratch,
                                                           int64_t first, uint32_t *__restrict__ status,
                                                           uint32_t *__restrict__ dbg) {
    const int slot = blockIdx.y, strip = blockIdx.x;
    if (slot >= n) return;
    const uint32_t cntv = slot_ptr<uint32_t>(scratch, L, slot, L.off_scnt)[strip];
    if (cntv > (uint32_t)L.slist_cap && threadIdx.x == 0) atomicOr(&status[first + slot], PAIR_ST_FALLBACK | 16u);   // reason 16: strip list overflow
    const uint32_t m = min(cntv, (uint32_t)L.slist_cap);
    const uint2 *slist = reinterpret_cast<const uint2 *>(slot_ptr<uint32_t>(scratch, L, slot, L.off_slist)) +
                         (size_t)strip * L.slist_cap;
    uint32_t *cnt = slot_ptr<uint32_t>(scratch, L, slot, L.off_cnt);
    uint16_t *cand = slot_ptr<uint16_t>(scratch, L, slot, L.off_cand);
    const int4 *rowpack = slot_ptr<int4>(scratch, L, slot, L.off_rowpack);
    const int32_t *lo_c = slot_ptr<int32_t>(scratch, L, slot, L.off_lo) + L.max_rows;
    const int32_t *w_c = slot_ptr<int32_t>(scratch, L, slot, L.off_w) + L.max_rows;
    unsigned total = 0u;
    for (uint32_t e = threadIdx.x; e < m; e += blockDim.x) {
        const uint2 rec = slist[e];
        if (rec.x == 0xffffffffu) continue;                   // unused tail of a lane's chunk
        ++total;
        const int i = rec.x & 0x3fff, j = (rec.x >> 14) & 0x3fff;
        const int z = (int)rec.y;
        const int4 rp = rowpack[i];                           // {aa_fix, rowLo - 2 EPS, rowW + 4 EPS, -}
        const int ar = z - rp.y, ac = z - (lo_c[j] - 2 * EPS);
        const bool zz = z < 2 * EPS;                          // near-zero item: always evaluated exactly
        const bool rz = (ar >= 0 && ar < rp.z) || zz;
        const bool cz = ac >= 0 && ac < w_c[j] + 4 * EPS;
        if (rz) {
            const unsigned p = atomicAdd(&cnt[i], 1u);
            if (p < CAND_CAP) cand[(size_t)i * CAND_CAP + p] = (uint16_t)(j | (ar < 2 * EPS ? 0x8000 : 0));
        }
        if (cz) {
            const int line = L.max_rows + j;
            const unsigned p = atomicAdd(&cnt[line], 1u);
            if (p < CAND_CAP) cand[(size_t)line * CAND_CAP + p] = (uint16_t)(i | (ac < 2 * EPS ? 0x8000 : 0));
        }
    }
    total = __reduce_add_sync(0xffffffffu, total);
    const int lane = threadIdx.x & 31;
    if (lane == 0 && total) atomicAdd(&dbg[24], total);
}

// ------------------------------------------------------------------------------------------------
// resolve: exact re-evaluation of candidate cells, exact order statistics, thresholds, bit patch
// ------------------------------------------------------------------------------------------------
// acc_f32prod with the float -> double widening done on the integer pipe (exact for normal non-negative
// floats; zero / subnormal products take the conversion instruction).  F2F.F64.F32 issues at 16 lanes/clk/SM
// (profiles/r1_ubench.md), so exact_item alternates the two forms and both pipes share the load.
__device__ __forceinline__ double acc_f32prod_alu(double acc, float a, float b) {
    const float p = __fmul_rn(a, b);
    const uint32_t bts = __float_as_uint(p);
    double w;
    if (bts - 0x00800000u < 0x7f000000u)                      // normal, positive, finite
        w = __hiloint2double((int)((bts >> 3) + 0x38000000u), (int)(bts << 29));
    else
        w = (double)p;
    return __dadd_rn(acc, w);
}

__device__ __forceinline__ float exact_item(const float *__restrict__ Q, const float *__restrict__ R, int i, int j,
                                            float aa, float bb) {
    // 9 consecutive frames of each side = 108 floats = 27 float4 (frames are 48 B, bases 16 B aligned)
    const float4 *a = reinterpret_cast<const float4 *>(Q + (int64_t)i * NBINS);
    const float4 *b = reinterpret_cast<const float4 *>(R + (int64_t)j * NBINS);
    double acc = 0.0;
#pragma unroll 3
    for (int t = 0; t < M9 * NBINS / 4; ++t) {
        const float4 u = __ldg(a + t), v = b[t];
        acc = acc_f32prod(acc, u.x, v.x); acc = acc_f32prod_alu(acc, u.y, v.y);
        acc = acc_f32prod(acc, u.z, v.z); acc = acc_f32prod_alu(acc, u.w, v.w);
    }
    return __fadd_rn(__fsub_rn(aa, __fmul_rn(2.f, (float)acc)), bb);
}

// 32 lines (rows, then columns) per CTA.  Phase 1 walks the CTA's candidates as one flat list, so every thread
// evaluates the same number of cells whatever the lines' counts are; phase 2 ranks with one 8-lane group per line.
__global__ void __launch_bounds__(256) fast_resolve_thr_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                               int64_t first, int n, FastLayout L,
                                                               char *__restrict__ scratch, int guard, double unit,
                                                               float *__restrict__ thr_q_all,
                                                               float *__restrict__ thr_r_all,
                                                               uint32_t *__restrict__ status) {
    __shared__ float s_item[32][CAND_CAP];
    __shared__ float s_sel[32][4];
    __shared__ int s_pref[33];                                // exclusive prefix of the lines' candidate counts
    __shared__ int s_nbelow[32];
    __shared__ int s_nan;
    const int slot = blockIdx.y;
    if (slot >= n) return;
    const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
    const int gline0 = blockIdx.x * 32;
    const int gline = gline0 + grp;                           // 0 .. Mx+Nx-1 (rows then columns)
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    const int Mx = h->Mx, Nx = h->Nx;
    const bool live = gline < Mx + Nx;
    const bool isrow = gline < Mx;
    const int idx = isrow ? gline : gline - Mx;
    const int line = isrow ? idx : L.max_rows + idx;
    const int64_t k = first + slot;
    const int q = pairs[2 * k];
    const float *Q = ts.frames + ts.offsets[q] * NBINS;
    const float *R = slot_ptr<float>(scratch, L, slot, L.off_rrot);
    const float *aaf = slot_ptr<float>(scratch, L, slot, L.off_aaf), *bbf = slot_ptr<float>(scratch, L, slot, L.off_bbf);
    const uint32_t *cnt_a = slot_ptr<uint32_t>(scratch, L, slot, L.off_cnt);
    const uint16_t *cand = slot_ptr<uint16_t>(scratch, L, slot, L.off_cand);
    float *candd = slot_ptr<float>(scratch, L, slot, L.off_candd);
    const unsigned gmask = 0xffu << (8 * ((threadIdx.x & 31) >> 3));
    if (threadIdx.x < 32) {                                   // warp 0: counts of the CTA's 32 lines and their prefix
        const int gl = gline0 + threadIdx.x;
        int c = 0;
        if (gl < Mx + Nx) c = (int)min(cnt_a[gl < Mx ? gl : L.max_rows + gl - Mx], (unsigned)CAND_CAP);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += v;
        }
        s_pref[threadIdx.x + 1] = incl;
        if (threadIdx.x == 0) { s_pref[0] = 0; s_nan = 0; }
        s_nbelow[threadIdx.x] = 0;
    }
    __syncthreads();
    const int total = s_pref[32];
    for (int wi = threadIdx.x; wi < total; wi += 256) {
        int g = 0;                                            // largest g with s_pref[g] <= wi
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1)
            if (s_pref[g + o] <= wi) g += o;
        const int p = wi - s_pref[g];
        const int gl = gline0 + g;
        const bool rw = gl < Mx;
        const int ix = rw ? gl : gl - Mx;
        const int ln = rw ? ix : L.max_rows + ix;
        const unsigned e = cand[(size_t)ln * CAND_CAP + p];
        const int other = e & 0x7fff;
        const int i = rw ? ix : other, j = rw ? other : ix;
        const float item = exact_item(Q, R, i, j, aaf[i], bbf[j]);
        const float d = __fsqrt_rn(item);
        if (d != d) s_nan = 1;
        candd[(size_t)ln * CAND_CAP + p] = d;
        s_item[g][p] = item;
        if (e >> 15) atomicAdd(&s_nbelow[g], 1);
    }
    __syncthreads();
    int cnt = 0;
    bool over = false;
    if (live) {
        const unsigned c = cnt_a[line];
        over = c > CAND_CAP;
        cnt = (int)min(c, (unsigned)CAND_CAP);
    }
    const int nbelow = s_nbelow[grp];
    const bool nan = (s_nan != 0) && (threadIdx.x == 0);
    if (!live) return;
    if (nan) atomicOr(&status[k], PAIR_ST_NAN);
    const int side = isrow ? 0 : 1;
    const int fk = h->fk[side], ck = h->ck[side];
    const int32_t lo = slot_ptr<int32_t>(scratch, L, slot, L.off_lo)[line];
    const int32_t w = slot_ptr<int32_t>(scratch, L, slot, L.off_w)[line];
    const int c0 = slot_ptr<int32_t>(scratch, L, slot, L.off_cb)[line] - nbelow;   // items below every candidate
    const int rfk = fk - c0, rck = ck - c0;
    unsigned bad = over ? 32u : 0u;                           // reason 32: more than CAND_CAP candidates on a line
    float thr = 0.f;
    if (!h->quirk[side]) {
        if (rfk < 0 || rck >= cnt || rfk > rck) bad |= 64u;   // reason 64: ranks not inside the candidate set
        // rank by counting (cnt <= 32): the candidate whose rank is rfk / rck publishes its item
        for (int p = sub; p < cnt; p += 8) {
            const float v = s_item[grp][p];
            int rank = 0;
            for (int p2 = 0; p2 < cnt; ++p2) {
                const float v2 = s_item[grp][p2];
                rank += (v2 < v) || (v2 == v && p2 < p);
            }
            if (rank == rfk) s_sel[grp][0] = v;
            if (rank == rck) s_sel[grp][1] = v;
        }
        __syncwarp(gmask);
        if (!bad) {
            const float ifk = s_sel[grp][0], ick = s_sel[grp][1];
            // every excluded cell below the candidate zone has exact item < (lo - EPS) units, every excluded cell
            // above it has exact item >= (lo + w + EPS) units: the selected order statistics must sit between
            const double lo_e = (double)(lo - EPS) * unit, hi_e = (double)(lo + w + EPS) * unit;
            if ((double)ifk < lo_e || (double)ick >= hi_e) bad |= 128u;   // reason 128: order statistic outside its zone
            const float sfk = __fsqrt_rn(ifk), sck = __fsqrt_rn(ick);
            const float kf = h->kf[side];
            const float fkf = floorf(kf), ckf = ceilf(kf);
            if (guard && fkf == ckf) thr = sfk;
            else thr = __fadd_rn(__fmul_rn(sfk, __fsub_rn(ckf, kf)), __fmul_rn(sck, __fsub_rn(kf, fkf)));
            // cells emitted as certainly-in have exact item < (lo - EPS): their d must be <= thr;
            // cells dropped as certainly-out have exact item >= (lo + w + EPS): their d must be > thr
            const double t2 = (double)thr * (double)thr;
            if (lo_e > 0.0 && t2 < lo_e * (1.0 + 1e-6)) bad |= 256u;   // reason 256/512: threshold not between the certain sets
            if (t2 >= hi_e * (1.0 - 1e-6)) bad |= 512u;
        }
    }
    if (sub == 0) {
        if (bad) atomicOr(&status[k], PAIR_ST_FALLBACK | bad);
        if (isrow) thr_q_all[(int64_t)slot * L.max_rows + idx] = thr;
        else thr_r_all[(int64_t)slot * L.max_cols + idx] = thr;
    }
}

__global__ void __launch_bounds__(256) fast_resolve_bits_kernel(int n, FastLayout L, char *__restrict__ scratch,
                                                                const float *__restrict__ thr_q_all,
                                                                const float *__restrict__ thr_r_all,
                                                                uint32_t *__restrict__ crp_all, int words,
                                                                int64_t crp_words) {
    const int slot = blockIdx.y;
    if (slot >= n) return;
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    const int Mx = h->Mx, Nx = h->Nx;
    const int gline = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;   // 16 lanes per line
    const int sub = threadIdx.x & 15;
    if (gline >= Mx + Nx) return;
    const bool isrow = gline < Mx;
    const int idx = isrow ? gline : gline - Mx;
    const int line = isrow ? idx : L.max_rows + idx;
    const int cnt = (int)min(slot_ptr<uint32_t>(scratch, L, slot, L.off_cnt)[line], (uint32_t)CAND_CAP);
    for (int p = sub; p < cnt; p += 16) {
    const unsigned e = slot_ptr<uint16_t>(scratch, L, slot, L.off_cand)[(size_t)line * CAND_CAP + p];
    const float d = slot_ptr<float>(scratch, L, slot, L.off_candd)[(size_t)line * CAND_CAP + p];
    const int other = e & 0x7fff;
    const int i = isrow ? idx : other, j = isrow ? other : idx;
    const float tq = thr_q_all[(int64_t)slot * L.max_rows + i], tr = thr_r_all[(int64_t)slot * L.max_cols + j];
    if ((__fsub_rn(tq, d) >= 0.f) && (__fsub_rn(tr, d) >= 0.f))
        atomicOr(crp_all + (int64_t)slot * crp_words + (int64_t)i * words + (j >> 5), 1u << (j & 31));
    }
}

__global__ void collect_fallback_kernel(const uint32_t *__restrict__ status, int64_t first, int n,
                                        int32_t *__restrict__ map, int32_t *__restrict__ count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (status[first + t] & PAIR_ST_FALLBACK) map[atomicAdd(count, 1)] = (int32_t)(first + t);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool k2_fast_supported(const acoss_params &p, const SlotGeom &g, const TrackSet &ts) {
    return p.m == M9 && p.tau == 1 && ts.fx_exp > -100 && ts.nonneg && g.max_rows < 16000 && g.max_cols < 16000 &&
           g.max_rows >= 2 && g.max_cols >= 2;
}

size_t k2_fast_slot_bytes(const SlotGeom &g, int max_frames) { return make_layout(g, max_frames).slot_bytes; }

int launch_k2_fast(const TrackSet &ts, const int32_t *pairs, const int32_t *oti, int64_t first, int n,
                   const acoss_params &p, const SlotGeom &g, void *scratch, size_t slot_bytes, uint32_t *crp,
                   float *thr_q, float *thr_r, uint32_t *status, uint32_t *dbg, cudaStream_t st, int64_t *launches,
                   cudaEvent_t emit_begin, cudaEvent_t emit_end) {
    if (n <= 0) return ACOSS_OK;
    constexpr int RC = RCV;
    const FastLayout L = make_layout(g, ts.max_frames);
    if (L.slot_bytes != slot_bytes) { acoss_set_error("fast path: scratch layout mismatch"); return ACOSS_E_INVALID; }
    char *base = (char *)scratch;
    const float qperc = (float)((double)(p.kappa * 100.f) / 100.);
    // fixed point: 2e < 2^fx_exp, unit = 2^(fx_exp - 23); magic = 2^fx_exp puts 2e in one binade
    const float magic = ldexpf(1.f, ts.fx_exp);
    const double unit = ldexp(1.0, ts.fx_exp - 23);
    const float fx_scale = ldexpf(1.f, 23 - ts.fx_exp);
    CUDA_TRY(cudaMemsetAsync(crp, 0, (size_t)n * g.crp_words * 4, st));
    fast_prep_kernel<<<n, 256, 0, st>>>(ts, pairs, oti, first, L, base, qperc, p.integer_guard, fx_scale);
    CUDA_TRY(cudaGetLastError());
    const int outw = Sweep<RC>::OUTW;
    const int strips_c = (g.max_cols + outw - 1) / outw, strips_r = (g.max_rows + outw - 1) / outw;
    const size_t smem = (size_t)WPC * (NBIN + 2) * (RC / 2) * 32 * 4;
    const size_t smem_sel = (size_t)(SBIN / 2) * SEL_THREADS * 4;
    static bool attr_done_dev[64] = {false};                  // function attributes are per device
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    bool &attr_done = attr_done_dev[dev & 63];
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(fast_hist_kernel<RC, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(fast_hist_kernel<RC, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(fast_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sel));
        attr_done = true;
    }
    const unsigned gc = (unsigned)(((int64_t)n * strips_c + WPC - 1) / WPC), gr = (unsigned)(((int64_t)n * strips_r + WPC - 1) / WPC);
    const int lines = g.max_rows + g.max_cols;
    // first brackets from every S-th diagonal (sampler + per-line selection), then up to three histogram
    // levels per orientation; a level returns immediately for strips whose lines are all done
    static const bool no_sample = getenv("ACOSS_K2_NO_SAMPLE") != nullptr;   // debugging: start from the whole item range
    if (!no_sample) {
        const int nu_max = ((g.max_rows - 1) >> L.slog) + ((g.max_cols - 1) >> L.slog) + 1;
        const int ngrp = (nu_max + SKD - 1) / SKD, nblk = (g.max_rows + SPOS - 1) / SPOS;
        const int64_t warps = (int64_t)n * ngrp * nblk;
        fast_sample_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(ts, pairs, first, n, L, base, ngrp, nblk, magic);
        CUDA_TRY(cudaGetLastError());
        fast_select_kernel<<<dim3((lines + SEL_THREADS - 1) / SEL_THREADS, n), SEL_THREADS, smem_sel, st>>>(n, L, base);
        CUDA_TRY(cudaGetLastError());
    }
    // two dense levels per orientation: the second sweeps only strips that still hold DENSE2_MIN_LIVE or more live
    // lines (short lines start from the whole item range and need it; ordinary strips skip it at once) and hands
    // every line still live to the sparse refinement
    fast_hist_kernel<RC, 0><<<gc, 32 * WPC, smem, st>>>(ts, pairs, first, n, L, base, strips_c, magic, status, dbg, 1, 0);
    fast_hist_kernel<RC, 1><<<gr, 32 * WPC, smem, st>>>(ts, pairs, first, n, L, base, strips_r, magic, status, dbg + 4, 1, 0);
    fast_hist_kernel<RC, 0><<<gc, 32 * WPC, smem, st>>>(ts, pairs, first, n, L, base, strips_c, magic, status, dbg + 12, DENSE2_MIN_LIVE, 1);
    fast_hist_kernel<RC, 1><<<gr, 32 * WPC, smem, st>>>(ts, pairs, first, n, L, base, strips_r, magic, status, dbg + 16, DENSE2_MIN_LIVE, 1);
    CUDA_TRY(cudaGetLastError());
    {
        const int64_t warps = (int64_t)n * 2 * (SPARSE_CAP / 32);
        fast_sparse_kernel<<<(unsigned)((warps + WPC - 1) / WPC), 32 * WPC, 0, st>>>(ts, pairs, first, n, L, base, magic, status, dbg + 8);
        CUDA_TRY(cudaGetLastError());
    }
    if (emit_begin) CUDA_TRY(cudaEventRecord(emit_begin, st));
    // 3 CTAs / SM (168 registers): 4 CTAs (128 registers, spills) measured 3 % slower
    fast_emit_kernel<RC><<<gc, 32 * WPC, 0, st>>>(ts, pairs, first, n, L, base, strips_c, magic, crp, g.words, g.crp_words);
    CUDA_TRY(cudaGetLastError());
    if (emit_end) CUDA_TRY(cudaEventRecord(emit_end, st));
    fast_scatter_kernel<<<dim3(strips_c, n), 128, 0, st>>>(n, L, base, first, status, dbg);
    CUDA_TRY(cudaGetLastError());
    fast_resolve_thr_kernel<<<dim3((lines + 31) / 32, n), 256, 0, st>>>(ts, pairs, first, n, L, base, p.integer_guard, unit,
                                                                       thr_q, thr_r, status);
    CUDA_TRY(cudaGetLastError());
    fast_resolve_bits_kernel<<<dim3((lines * 16 + 255) / 256, n), 256, 0, st>>>(n, L, base, thr_q, thr_r, crp, g.words,
                                                                                      g.crp_words);
    CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 16;
    return ACOSS_OK;
}

int k2_fast_collect_fallback(const uint32_t *status, int64_t first, int n, int32_t *map_dev, int32_t *count_dev,
                             int *count_host, cudaStream_t st) {
    *count_host = 0;
    if (n <= 0) return ACOSS_OK;
    CUDA_TRY(cudaMemsetAsync(count_dev, 0, 4, st));
    collect_fallback_kernel<<<(n + 255) / 256, 256, 0, st>>>(status, first, n, map_dev, count_dev);
    CUDA_TRY(cudaGetLastError());
    int32_t c = 0;
    CUDA_TRY(cudaMemcpyAsync(&c, count_dev, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *count_host = c;
    return ACOSS_OK;
}
