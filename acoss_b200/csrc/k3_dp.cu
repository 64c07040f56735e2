// K3 — Qmax / Dmax / constrained Smith-Waterman local-alignment DP over bit-packed binary matrices.
//
// Replaces (file:line under /root/reference, SURVEY.md App. A6 for the essentia part):
//   essentia CoverSongSimilarity(alignmentType='serra09', distanceType='symmetric')   rqa_serra09.py:64,67
//   essentia CoverSongSimilarity(alignmentType='chen17',  distanceType='symmetric')   latefusion_chen.py:69-73
//   smith_waterman_constrained(B)                      acoss/algorithms/utils/alignment_tools.py:26-46
//
// Both recurrences read only rows i-1 and i-2 (predecessors (i-1,j-1), (i-2,j-1), (i-1,j-2)), so a
// whole row updates in parallel and rows are the only serial axis.  One warp owns one pair and
// sweeps rows; each lane keeps the previous two DP rows of its contiguous column chunk in
// registers, and only three halo values cross lanes per row (warp shuffles).  Many pairs per CTA.
//
// Packed path (dp_packed_kernel): exact integer DP in packed signed 16-bit pairs using the DPX
// instructions VIMNMX3.S16x2 / VIADDMNMX.S16x2.RELU:
//   Qmax with gamma_o = gamma_e = 0.5 (essentia defaults; acoss passes none): scores scaled x2,
//        Q2 = relu(max3 + (c ? +2 : -1)),           bounded by 2*min(M',N')  -> int16 up to 16383
//   SW-constrained: scores scaled x10 (match +10, mismatch -10, gap -7):
//        S  = relu(max3(S'') + (b ? +3 : -17)),  S'' = S + 7 b   (S'' = S + delta + 7 >= 0),
//        bounded by 10*min(M,N) + 7                  -> int16 up to 3275
// Scalar path (dp_scalar_kernel): float32, any gamma_o / gamma_e, any size; same operations per cell
// as the oracle so results are bit-identical floats.
//
// Register layout of the packed path: a lane owns G 32-column groups; packed register t of a group
// holds columns (t, t+16) of the group in its (low, high) halves, so "column - 1" of register t is
// register t-1 and one funnel shift + one mask extracts both weights from the CRP word.
#include "common.cuh"

#define MODE_QMAX 0
#define MODE_SW 1

__device__ __forceinline__ uint32_t lohi(uint32_t prev, uint32_t own) {
    // result.low = prev.high, result.high = own.low
    return __byte_perm(prev, own, 0x5432);
}

template <int MODE>
struct PackedW {
    // weights per half: hit / miss added to max3, and the stored-value bias for hits
    static constexpr uint32_t WPOS = (MODE == MODE_QMAX) ? 0x00020002u : 0x00030003u;   // +2 | +3
    static constexpr uint32_t WNEG = (MODE == MODE_QMAX) ? 0xffffffffu : 0xffefffefu;   // -1 | -17
    static constexpr uint32_t BIAS = (MODE == MODE_QMAX) ? 0u : 7u;                      // S'' = S + 7 b
};

// One DP row for one 32-column group.  A = row u-1, B = row u-2 (both "stored" values X); new row
// is written over B (descending t keeps B[t-1] alive until used).  Returns via best the max of S.
template <int MODE>
__device__ __forceinline__ void group_row(uint32_t (&A)[16], uint32_t (&B)[16], uint32_t hA1, uint32_t hA2,
                                          uint32_t hB1, uint32_t bits, uint32_t &best) {
#pragma unroll
    for (int t = 15; t >= 0; --t) {
        const uint32_t a1 = (t >= 1) ? A[t - 1] : hA1;
        const uint32_t b1 = (t >= 1) ? B[t - 1] : hB1;
        const uint32_t a2 = (t >= 2) ? A[t - 2] : (t == 1 ? hA1 : hA2);
        const uint32_t v = (bits >> t) & 0x00010001u;
        const uint32_t msk = v * 0xffffu;
        const uint32_t w = (msk & PackedW<MODE>::WPOS) | (~msk & PackedW<MODE>::WNEG);
        const uint32_t m3 = __vimax3_s16x2(a1, b1, a2);
        const uint32_t s = __viaddmax_s16x2_relu(m3, w, 0u);
        best = __vmaxs2(best, s);
        B[t] = (MODE == MODE_QMAX) ? s : (s + v * PackedW<MODE>::BIAS);   // no carry: halves stay < 2^15
    }
}

template <int G, int MODE>
__global__ void __launch_bounds__(128) dp_packed_kernel(const uint32_t *__restrict__ bits_all, int64_t slot_words,
                                                        int wpr, const int32_t *__restrict__ rows_a,
                                                        const int32_t *__restrict__ cols_a, int n,
                                                        float *__restrict__ scores, uint32_t *__restrict__ halo_all,
                                                        int64_t halo_pitch) {
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= n) return;
    const int lane = threadIdx.x & 31;
    const int R = rows_a[pair], C = cols_a[pair];
    if (R < 3 || C < 3) {
        if (lane == 0) scores[pair] = 0.f;
        return;
    }
    const uint32_t *bits = bits_all + (int64_t)pair * slot_words;
    constexpr int W = 1024 * G;                  // DP columns per strip
    const int nstrips = (C - 2 + W - 1) / W;
    uint32_t *halo0 = halo_all ? halo_all + (int64_t)pair * 2 * halo_pitch : nullptr;
    uint32_t best = 0u;
    constexpr uint32_t BIAS = PackedW<MODE>::BIAS;

    for (int s = 0; s < nstrips; ++s) {
        const int wbase = s * 32 * G + lane * G;            // first CRP word of this lane's chunk
        const uint32_t *halo_in = halo0 ? halo0 + (int64_t)(s & 1) * halo_pitch : nullptr;
        uint32_t *halo_out = halo0 ? halo0 + (int64_t)((s + 1) & 1) * halo_pitch : nullptr;
        const bool write_halo = (s + 1 < nstrips);

        uint32_t Wd[G + 1];
        auto load_row = [&](int u, uint32_t (&dst)[G + 1]) {
            const uint32_t *row = bits + (int64_t)u * wpr;
#pragma unroll
            for (int g = 0; g <= G; ++g) {
                const int w = wbase + g;
                dst[g] = (w < wpr) ? __ldg(row + w) : 0u;
            }
        };
        // stored halo word of row u for lane 0: (X[u][c0-2] | X[u][c0-1] << 16)
        auto halo_word = [&](int u, uint32_t w0) -> uint32_t {
            if (s == 0 || u < 2) return ((w0 & 1u) | ((w0 & 2u) << 15)) * BIAS;
            return halo_in[u];
        };

        uint32_t A[G][16], B[G][16];
        uint32_t hprev1, hprev2;
        // rows 0 and 1: X = BIAS * b
        {
            uint32_t W0[G + 1], W1[G + 1];
            load_row(0, W0);
            load_row(1, W1);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint32_t b0 = __funnelshift_r(W0[g], W0[g + 1], 2), b1 = __funnelshift_r(W1[g], W1[g + 1], 2);
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    B[g][t] = ((b0 >> t) & 0x00010001u) * BIAS;
                    A[g][t] = ((b1 >> t) & 0x00010001u) * BIAS;
                }
            }
            hprev2 = halo_word(0, W0[0]);
            hprev1 = halo_word(1, W1[0]);
        }
        load_row(2, Wd);

        auto row_step = [&](int u, uint32_t (&Ar)[G][16], uint32_t (&Br)[G][16]) {
            // halos (taken before any overwrite of Br)
            uint32_t nA15 = __shfl_up_sync(0xffffffffu, Ar[G - 1][15], 1);
            uint32_t nA14 = __shfl_up_sync(0xffffffffu, Ar[G - 1][14], 1);
            uint32_t nB15 = __shfl_up_sync(0xffffffffu, Br[G - 1][15], 1);
            if (lane == 0) { nA15 = hprev1; nA14 = hprev1 << 16; nB15 = hprev2; }
            uint32_t hA1[G], hA2[G], hB1[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint32_t pA15 = (g == 0) ? nA15 : Ar[g - 1][15];
                const uint32_t pA14 = (g == 0) ? nA14 : Ar[g - 1][14];
                const uint32_t pB15 = (g == 0) ? nB15 : Br[g - 1][15];
                hA1[g] = lohi(pA15, Ar[g][15]);
                hA2[g] = lohi(pA14, Ar[g][14]);
                hB1[g] = lohi(pB15, Br[g][15]);
            }
            uint32_t cur[G + 1];
#pragma unroll
            for (int g = 0; g <= G; ++g) cur[g] = Wd[g];
            if (u + 1 < R) load_row(u + 1, Wd);                 // prefetch next row's CRP words
            const uint32_t hnew = (lane == 0) ? halo_word(u, cur[0]) : 0u;
#pragma unroll
            for (int g = 0; g < G; ++g)
                group_row<MODE>(Ar[g], Br[g], hA1[g], hA2[g], hB1[g], __funnelshift_r(cur[g], cur[g + 1], 2), best);
            if (write_halo && lane == 31) halo_out[u] = __byte_perm(Br[G - 1][14], Br[G - 1][15], 0x7632);
            hprev2 = hprev1;
            hprev1 = hnew;
        };

        int u = 2;
        for (; u + 1 < R; u += 2) {
            row_step(u, A, B);       // new row -> B
            row_step(u + 1, B, A);   // new row -> A
        }
        if (u < R) row_step(u, A, B);
        __syncwarp();
    }
    // best holds two int16 maxima
    int bm = max((int)(short)(best & 0xffffu), (int)(short)(best >> 16));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) bm = max(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    if (lane == 0) scores[pair] = (MODE == MODE_QMAX) ? (float)bm * 0.5f : (float)((double)bm / 10.0);
}

// ------------------------------------------------------------------------------------------------
// Scalar float32 path: any gamma, any size.  Lane owns CW contiguous DP columns.
//   V = value read by "hit" cells, P = value read by "miss" cells.
//   Qmax: new = c ? max3(V)+1 : relu(max3(P)) ; V = new ; P = new - (c ? go : ge)
//   SW (x10 integers held exactly in float32): new = relu(max3(P) + (c ? 3 : -17)) ; P = new + 7c
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) dp_scalar_kernel(const uint32_t *__restrict__ bits_all, int64_t slot_words,
                                                        int wpr, const int32_t *__restrict__ rows_a,
                                                        const int32_t *__restrict__ cols_a, int n, float go, float ge,
                                                        float *__restrict__ scores, float4 *__restrict__ halo_all,
                                                        int64_t halo_pitch) {
    constexpr int CW = 8, W = 32 * CW;
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= n) return;
    const int lane = threadIdx.x & 31;
    const int R = rows_a[pair], C = cols_a[pair];
    if (R < 3 || C < 3) {
        if (lane == 0) scores[pair] = 0.f;
        return;
    }
    const uint32_t *bits = bits_all + (int64_t)pair * slot_words;
    const int nstrips = (C - 2 + W - 1) / W;
    float4 *halo0 = halo_all + (int64_t)pair * 2 * halo_pitch;
    float best = 0.f;
    auto bit_at = [&](int u, int c) -> int {        // CRP bit (u, c), 0 outside the matrix
        if (c < 0 || c >= C) return 0;
        return (__ldg(bits + (int64_t)u * wpr + (c >> 5)) >> (c & 31)) & 1;
    };
    auto vinit = [&](int c) -> float { return 0.f; };
    auto pinit = [&](int c) -> float {              // boundary P from the bit
        if (MODE == MODE_QMAX) return -(c ? go : ge);
        return c ? 7.f : 0.f;
    };
    for (int s = 0; s < nstrips; ++s) {
        const int c0 = 2 + s * W + lane * CW;       // CRP column of this lane's first DP column
        const float4 *halo_in = halo0 + (int64_t)(s & 1) * halo_pitch;
        float4 *halo_out = halo0 + (int64_t)((s + 1) & 1) * halo_pitch;
        const bool write_halo = (s + 1 < nstrips);
        float Va[CW], Pa[CW], Vb[CW], Pb[CW];       // a = row u-1, b = row u-2
#pragma unroll
        for (int k = 0; k < CW; ++k) {
            Va[k] = vinit(0); Vb[k] = vinit(0);
            Pa[k] = pinit(bit_at(1, c0 + k));
            Pb[k] = pinit(bit_at(0, c0 + k));
        }
        // halo of the two columns left of the chunk for rows u-1 (h1*) and u-2 (h2*): x = col-2, y = col-1
        auto boundary = [&](int u) -> float4 {      // (V[c-2], V[c-1], P[c-2], P[c-1]) at strip start
            const int cb = 2 + s * W;
            if (s == 0 || u < 2) return make_float4(0.f, 0.f, pinit(bit_at(u, cb - 2)), pinit(bit_at(u, cb - 1)));
            return halo_in[u];
        };
        float4 hp1 = boundary(1), hp2 = boundary(0);
        for (int u = 2; u < R; ++u) {
            // left neighbours' last two columns of rows u-1 / u-2
            float lV1 = __shfl_up_sync(0xffffffffu, Va[CW - 1], 1), lV2 = __shfl_up_sync(0xffffffffu, Va[CW - 2], 1);
            float lP1 = __shfl_up_sync(0xffffffffu, Pa[CW - 1], 1), lP2 = __shfl_up_sync(0xffffffffu, Pa[CW - 2], 1);
            float lVb = __shfl_up_sync(0xffffffffu, Vb[CW - 1], 1), lPb = __shfl_up_sync(0xffffffffu, Pb[CW - 1], 1);
            if (lane == 0) { lV2 = hp1.x; lV1 = hp1.y; lP2 = hp1.z; lP1 = hp1.w; lVb = hp2.y; lPb = hp2.w; }
            float Vn[CW], Pn[CW];
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                const float v1 = (k >= 1) ? Va[k - 1] : lV1, p1 = (k >= 1) ? Pa[k - 1] : lP1;
                const float vb = (k >= 1) ? Vb[k - 1] : lVb, pb = (k >= 1) ? Pb[k - 1] : lPb;
                const float v2 = (k >= 2) ? Va[k - 2] : (k == 1 ? lV1 : lV2);
                const float p2 = (k >= 2) ? Pa[k - 2] : (k == 1 ? lP1 : lP2);
                const int c = bit_at(u, c0 + k);
                float nv;
                if (MODE == MODE_QMAX) {
                    const float hit = __fadd_rn(fmaxf(fmaxf(v1, vb), v2), 1.f);
                    const float miss = fmaxf(fmaxf(p1, pb), fmaxf(p2, 0.f));
                    nv = c ? hit : miss;
                    Vn[k] = nv;
                    Pn[k] = __fsub_rn(nv, c ? go : ge);
                } else {
                    nv = fmaxf(__fadd_rn(fmaxf(fmaxf(p1, pb), p2), c ? 3.f : -17.f), 0.f);
                    Vn[k] = nv;
                    Pn[k] = nv + (c ? 7.f : 0.f);
                }
                if (c0 + k < C) best = fmaxf(best, nv);
            }
            const float4 hnew = (lane == 0) ? boundary(u) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (write_halo && lane == 31) halo_out[u] = make_float4(Vn[CW - 2], Vn[CW - 1], Pn[CW - 2], Pn[CW - 1]);
#pragma unroll
            for (int k = 0; k < CW; ++k) { Vb[k] = Va[k]; Pb[k] = Pa[k]; Va[k] = Vn[k]; Pa[k] = Pn[k]; }
            hp2 = hp1;
            hp1 = hnew;
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) scores[pair] = (MODE == MODE_QMAX) ? best : (float)((double)best / 10.0);
}

// ------------------------------------------------------------------------------------------------
// Dmax (essentia alignmentType='chen17'; ChenFusion's second score, latefusion_chen.py:69-73), float32,
// any gamma.  Five predecessors (u-1,c-1), (u-2,c-1), (u-1,c-2), (u-3,c-1), (u-1,c-3); with BONUS (F10) the
// skipped-cell predecessors gain Chen's bridging bits:
//   P2 += b(u-1,c)   P3 += b(u,c-1)   P4 += b(u-2,c) + b(u-1,c)   P5 += b(u,c-2) + b(u,c-1)
// No same-row dependency either: the row sweep of the other recurrences applies with three previous rows
// in registers.  A lane owns CW contiguous columns and keeps, per row, a bit window of columns
// [c0-3, c0+CW) in one register; five halo values cross lanes per row; strips hand the last three columns
// of every row to the next strip.  Same operations per cell as the oracle: bit-identical floats.
// ------------------------------------------------------------------------------------------------
template <int BONUS>
__global__ void __launch_bounds__(128) dp_dmax_kernel(const uint32_t *__restrict__ bits_all, int64_t slot_words,
                                                      int wpr, const int32_t *__restrict__ rows_a,
                                                      const int32_t *__restrict__ cols_a, int n, float go, float ge,
                                                      float *__restrict__ scores, float4 *__restrict__ halo_all,
                                                      int64_t halo_pitch) {
    constexpr int CW = 8, W = 32 * CW;
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= n) return;
    const int lane = threadIdx.x & 31;
    const int R = rows_a[pair], C = cols_a[pair];
    if (R < 4 || C < 4) {
        if (lane == 0) scores[pair] = 0.f;
        return;
    }
    const uint32_t *bits = bits_all + (int64_t)pair * slot_words;
    const int nstrips = (C - 3 + W - 1) / W;
    float4 *halo0 = halo_all + (int64_t)pair * 2 * halo_pitch;
    float best = 0.f;
    for (int s = 0; s < nstrips; ++s) {
        const int c0 = 3 + s * W + lane * CW;       // CRP column of this lane's first DP column
        const float4 *halo_in = halo0 + (int64_t)(s & 1) * halo_pitch;
        float4 *halo_out = halo0 + (int64_t)((s + 1) & 1) * halo_pitch;
        const bool write_halo = (s + 1 < nstrips);
        // bit k+3+off of a window = CRP bit (row, c0 + k + off); columns >= C are pad zeros
        auto window = [&](int u) -> uint32_t {
            const uint32_t *row = bits + (int64_t)u * wpr;
            const int cw = c0 - 3, w = cw >> 5;
            const uint32_t lo = (w < wpr) ? __ldg(row + w) : 0u, hi = (w + 1 < wpr) ? __ldg(row + w + 1) : 0u;
            return __funnelshift_r(lo, hi, cw & 31);
        };
        float D1[CW], D2[CW], D3[CW];               // rows u-1, u-2, u-3
#pragma unroll
        for (int k = 0; k < CW; ++k) { D1[k] = 0.f; D2[k] = 0.f; D3[k] = 0.f; }
        uint32_t W1 = window(2), W2 = window(1), W3 = window(0);
        // (D[c-3], D[c-2], D[c-1]) left of the strip for rows u-1, u-2, u-3 (lane 0 only); rows < 3 are zero
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 h1 = zero4, h2 = zero4, h3 = zero4;
        uint32_t Wn = window(3);
        for (int u = 3; u < R; ++u) {
            const uint32_t W0 = Wn;
            if (u + 1 < R) Wn = window(u + 1);      // prefetch the next row's bits
            float l1a = __shfl_up_sync(0xffffffffu, D1[CW - 1], 1), l1b = __shfl_up_sync(0xffffffffu, D1[CW - 2], 1);
            float l1c = __shfl_up_sync(0xffffffffu, D1[CW - 3], 1);
            float l2 = __shfl_up_sync(0xffffffffu, D2[CW - 1], 1), l3 = __shfl_up_sync(0xffffffffu, D3[CW - 1], 1);
            if (lane == 0) { l1a = h1.z; l1b = h1.y; l1c = h1.x; l2 = h2.z; l3 = h3.z; }
            float Dn[CW];
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                const float d1m1 = (k >= 1) ? D1[k - 1] : l1a;
                const float d1m2 = (k >= 2) ? D1[k - 2] : (k == 1 ? l1a : l1b);
                const float d1m3 = (k >= 3) ? D1[k - 3] : (k == 2 ? l1a : (k == 1 ? l1b : l1c));
                const float d2m1 = (k >= 1) ? D2[k - 1] : l2;
                const float d3m1 = (k >= 1) ? D3[k - 1] : l3;
                auto bit = [&](uint32_t wv, int off) -> uint32_t { return (wv >> (k + 3 + off)) & 1u; };
                float P1 = d1m1, P2 = d2m1, P3 = d1m2, P4 = d3m1, P5 = d1m3;
                if (BONUS) {
                    P2 = __fadd_rn(P2, (float)bit(W1, 0));
                    P3 = __fadd_rn(P3, (float)bit(W0, -1));
                    P4 = __fadd_rn(__fadd_rn(P4, (float)bit(W2, 0)), (float)bit(W1, 0));
                    P5 = __fadd_rn(__fadd_rn(P5, (float)bit(W0, -2)), (float)bit(W0, -1));
                }
                float v;
                if (bit(W0, 0)) {
                    v = __fadd_rn(fmaxf(fmaxf(fmaxf(P1, P2), fmaxf(P3, P4)), P5), 1.f);
                } else {
                    P1 = __fsub_rn(P1, bit(W1, -1) ? go : ge);
                    P2 = __fsub_rn(P2, bit(W2, -1) ? go : ge);
                    P3 = __fsub_rn(P3, bit(W1, -2) ? go : ge);
                    P4 = __fsub_rn(P4, bit(W3, -1) ? go : ge);
                    P5 = __fsub_rn(P5, bit(W1, -3) ? go : ge);
                    v = fmaxf(fmaxf(fmaxf(P1, P2), fmaxf(P3, P4)), fmaxf(P5, 0.f));
                }
                Dn[k] = v;
                if (c0 + k < C) best = fmaxf(best, v);
            }
            const float4 hnew = (lane == 0 && s > 0) ? halo_in[u] : zero4;
            if (write_halo && lane == 31) halo_out[u] = make_float4(Dn[CW - 3], Dn[CW - 2], Dn[CW - 1], 0.f);
#pragma unroll
            for (int k = 0; k < CW; ++k) { D3[k] = D2[k]; D2[k] = D1[k]; D1[k] = Dn[k]; }
            W3 = W2; W2 = W1; W1 = W0;
            h3 = h2; h2 = h1; h1 = hnew;
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) scores[pair] = best;
}

// ------------------------------------------------------------------------------------------------
// Dmax, packed: exact integer DP in signed 16-bit pairs for essentia's default gamma_o = gamma_e = 0.5 with
// Chen's bridging terms, scores scaled x2:
//   D2 = relu(max5 + (c ? +2 : -1)),  max5 over A[c-1], B[c-1] + 2 b(u-1,c), A[c-2] + 2 b(u,c-1),
//                                      C[c-1] + 2 (b(u-2,c) + b(u-1,c)), A[c-3] + 2 (b(u,c-2) + b(u,c-1))
// (A, B, C = rows u-1, u-2, u-3).  Every move gains at most rows + columns advanced - 1, so D <= R + C and
// D2 fits int16 while 2 (R + C) <= 32000.  Same register layout as dp_packed_kernel (register t of a
// 32-column group = columns t, t+16); the new row is written over the oldest one (C), t descending.
// Cells right of the matrix (pad bits 0) never exceed the largest real cell: by induction a pad cell is
// relu(max5 - 1) with zero bridging bits in its own column, and every candidate it sees is matched or beaten
// by the real cell left of it (a hit there adds 2) -- so no column mask is needed for the maximum.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmax_group_row(const uint32_t (&A)[16], const uint32_t (&B)[16], uint32_t (&C)[16],
                                               uint32_t hA1, uint32_t hA2, uint32_t hA3, uint32_t hB1, uint32_t hC1,
                                               uint32_t w0, uint32_t w1, uint32_t w2, uint32_t pw0, uint32_t ppw0,
                                               uint32_t &best) {
    constexpr uint32_t M = 0x00010001u;
#pragma unroll
    for (int t = 15; t >= 0; --t) {
        const uint32_t a1 = (t >= 1) ? A[t - 1] : hA1;
        const uint32_t a2 = (t >= 2) ? A[t - 2] : (t == 1 ? hA1 : hA2);
        const uint32_t a3 = (t >= 3) ? A[t - 3] : (t == 2 ? hA1 : (t == 1 ? hA2 : hA3));
        const uint32_t b1 = (t >= 1) ? B[t - 1] : hB1;
        const uint32_t c1 = (t >= 1) ? C[t - 1] : hC1;
        const uint32_t v0 = (w0 >> t) & M;                     // b(u, c)
        const uint32_t xu1 = (w1 >> t) & M, xu2 = (w2 >> t) & M;   // b(u, c-1), b(u, c-2)
        const uint32_t xp = (pw0 >> t) & M, xpp = (ppw0 >> t) & M; // b(u-1, c), b(u-2, c)
        const uint32_t c2 = b1 + 2u * xp;                      // halves stay < 2^15: no carry between them
        const uint32_t c3 = a2 + 2u * xu1;
        const uint32_t c4 = c1 + 2u * (xpp + xp);
        const uint32_t c5 = a3 + 2u * (xu2 + xu1);
        const uint32_t m5 = __vimax3_s16x2(__vimax3_s16x2(a1, c2, c3), c4, c5);
        const uint32_t msk = v0 * 0xffffu;
        const uint32_t w = (msk & 0x00020002u) | ~msk;         // +2 on a hit, -1 otherwise
        const uint32_t sres = __viaddmax_s16x2_relu(m5, w, 0u);
        best = __vmaxs2(best, sres);
        C[t] = sres;
    }
}

template <int G>
__global__ void __launch_bounds__(128) dp_dmax_packed_kernel(const uint32_t *__restrict__ bits_all, int64_t slot_words,
                                                             int wpr, const int32_t *__restrict__ rows_a,
                                                             const int32_t *__restrict__ cols_a, int n,
                                                             float *__restrict__ scores, uint2 *__restrict__ halo_all,
                                                             int64_t halo_pitch) {
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= n) return;
    const int lane = threadIdx.x & 31;
    const int R = rows_a[pair], C = cols_a[pair];
    if (R < 4 || C < 4) {
        if (lane == 0) scores[pair] = 0.f;
        return;
    }
    const uint32_t *bits = bits_all + (int64_t)pair * slot_words;
    constexpr int W = 1024 * G;                  // DP columns per strip (DP column d = CRP column d + 3)
    const int nstrips = (C - 3 + W - 1) / W;
    uint2 *halo0 = halo_all + (int64_t)pair * 2 * halo_pitch;
    uint32_t best = 0u;
    for (int s = 0; s < nstrips; ++s) {
        const int wbase = s * 32 * G + lane * G;            // first CRP word of this lane's chunk
        const uint2 *halo_in = halo0 + (int64_t)(s & 1) * halo_pitch;
        uint2 *halo_out = halo0 + (int64_t)((s + 1) & 1) * halo_pitch;
        const bool write_halo = (s + 1 < nstrips);
        uint32_t Wd[G + 1];
        auto load_row = [&](int u, uint32_t (&dst)[G + 1]) {
            const uint32_t *row = bits + (int64_t)u * wpr;
#pragma unroll
            for (int g = 0; g <= G; ++g) {
                const int w = wbase + g;
                dst[g] = (w < wpr) ? __ldg(row + w) : 0u;
            }
        };
        uint32_t A[G][16], B[G][16], Cc[G][16];
        uint32_t pw0[G], ppw0[G];
        {
            uint32_t T1[G + 1], T2[G + 1];
            load_row(1, T1);
            load_row(2, T2);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                ppw0[g] = __funnelshift_r(T1[g], T1[g + 1], 3);
                pw0[g] = __funnelshift_r(T2[g], T2[g + 1], 3);
#pragma unroll
                for (int t = 0; t < 16; ++t) { A[g][t] = 0u; B[g][t] = 0u; Cc[g][t] = 0u; }
            }
        }
        // lane 0: (D2[c0-2] | D2[c0-1] << 16, D2[c0-3]) of rows u-1, u-2, u-3 at the strip's left edge; rows < 3 are zero
        uint2 h1 = make_uint2(0u, 0u), h2 = h1, h3 = h1;
        load_row(3, Wd);

        auto row_step = [&](int u, uint32_t (&Ar)[G][16], uint32_t (&Br)[G][16], uint32_t (&Cr)[G][16]) {
            uint32_t nA15 = __shfl_up_sync(0xffffffffu, Ar[G - 1][15], 1);
            uint32_t nA14 = __shfl_up_sync(0xffffffffu, Ar[G - 1][14], 1);
            uint32_t nA13 = __shfl_up_sync(0xffffffffu, Ar[G - 1][13], 1);
            uint32_t nB15 = __shfl_up_sync(0xffffffffu, Br[G - 1][15], 1);
            uint32_t nC15 = __shfl_up_sync(0xffffffffu, Cr[G - 1][15], 1);
            if (lane == 0) { nA15 = h1.x; nA14 = h1.x << 16; nA13 = h1.y << 16; nB15 = h2.x; nC15 = h3.x; }
            uint32_t hA1[G], hA2[G], hA3[G], hB1[G], hC1[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                hA1[g] = lohi((g == 0) ? nA15 : Ar[g - 1][15], Ar[g][15]);
                hA2[g] = lohi((g == 0) ? nA14 : Ar[g - 1][14], Ar[g][14]);
                hA3[g] = lohi((g == 0) ? nA13 : Ar[g - 1][13], Ar[g][13]);
                hB1[g] = lohi((g == 0) ? nB15 : Br[g - 1][15], Br[g][15]);
                hC1[g] = lohi((g == 0) ? nC15 : Cr[g - 1][15], Cr[g][15]);
            }
            uint32_t cur[G + 1];
#pragma unroll
            for (int g = 0; g <= G; ++g) cur[g] = Wd[g];
            if (u + 1 < R) load_row(u + 1, Wd);                 // prefetch next row's CRP words
            const uint2 hnew = (lane == 0 && s > 0) ? halo_in[u] : make_uint2(0u, 0u);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const uint32_t w0 = __funnelshift_r(cur[g], cur[g + 1], 3), w1 = __funnelshift_r(cur[g], cur[g + 1], 2),
                               w2 = __funnelshift_r(cur[g], cur[g + 1], 1);
                dmax_group_row(Ar[g], Br[g], Cr[g], hA1[g], hA2[g], hA3[g], hB1[g], hC1[g], w0, w1, w2, pw0[g], ppw0[g], best);
                ppw0[g] = pw0[g];
                pw0[g] = w0;
            }
            if (write_halo && lane == 31)
                halo_out[u] = make_uint2(__byte_perm(Cr[G - 1][14], Cr[G - 1][15], 0x7632), Cr[G - 1][13] >> 16);
            h3 = h2; h2 = h1; h1 = hnew;
        };

        int u = 3;
        for (; u + 2 < R; u += 3) {
            row_step(u, A, B, Cc);          // new row -> Cc
            row_step(u + 1, Cc, A, B);      // new row -> B
            row_step(u + 2, B, Cc, A);      // new row -> A
        }
        if (u < R) { row_step(u, A, B, Cc); ++u; }
        if (u < R) { row_step(u, Cc, A, B); ++u; }
        __syncwarp();
    }
    int bm = max((int)(short)(best & 0xffffu), (int)(short)(best >> 16));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) bm = max(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    if (lane == 0) scores[pair] = (float)bm * 0.5f;
}

// ------------------------------------------------------------------------------------------------
// helpers: pair geometry, byte-matrix packing
// ------------------------------------------------------------------------------------------------
__global__ void pair_geometry_kernel(TrackSet ts, const int32_t *__restrict__ pairs, int64_t first, int n, int incr,
                                     int32_t *__restrict__ rows, int32_t *__restrict__ cols) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int q = pairs[2 * (first + k)], r = pairs[2 * (first + k) + 1];
    rows[k] = (int)(ts.offsets[q + 1] - ts.offsets[q]) - incr;
    cols[k] = (int)(ts.offsets[r + 1] - ts.offsets[r]) - incr;
}

// One warp per (matrix row, 32-column word).  bit = (B > 0); cells the reference validates
// (alignment_tools.py:17-23 reads B[i-1][j-1], i in 3..M-1, j in 3..N-1; essentia checks every cell)
// must be exactly 0 or 1, otherwise *flag is raised.
__global__ void __launch_bounds__(256) pack_bytes_kernel(const uint8_t *__restrict__ mats,
                                                         const int64_t *__restrict__ offsets,
                                                         const int32_t *__restrict__ shapes, int mode,
                                                         uint32_t *__restrict__ bits_all, int64_t slot_words, int wpr,
                                                         int32_t *__restrict__ rows, int32_t *__restrict__ cols,
                                                         uint32_t *__restrict__ flag) {
    const int k = blockIdx.z, i = blockIdx.y;
    const int M = shapes[2 * k], N = shapes[2 * k + 1];
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i == 0 && w == 0 && lane == 0) {
        // DP matrix: SW drops the last row/column of B (never read by the reference)
        rows[k] = (mode == ACOSS_ALIGN_SW) ? M - 1 : M;
        cols[k] = (mode == ACOSS_ALIGN_SW) ? N - 1 : N;
    }
    if (i >= M || w >= wpr) return;
    const int j = w * 32 + lane;
    const int jmax = (mode == ACOSS_ALIGN_SW) ? N - 1 : N;     // columns >= jmax are masked to 0
    bool bit = false;
    if (j < jmax) {
        const uint8_t v = mats[offsets[k] + (int64_t)i * N + j];
        bit = v > 0;
        const bool checked = (mode == ACOSS_ALIGN_SW) ? (i >= 2 && i <= M - 2 && j >= 2 && j <= N - 2) : true;
        if (checked && v > 1) atomicOr(flag, 1u);
    }
    const unsigned word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) bits_all[(int64_t)k * slot_words + (int64_t)i * wpr + w] = word;
}

int launch_pair_geometry(const TrackSet &ts, const int32_t *pairs, int64_t first, int n, int incr, int32_t *rows,
                         int32_t *cols, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    pair_geometry_kernel<<<(n + 255) / 256, 256, 0, st>>>(ts, pairs, first, n, incr, rows, cols);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

__global__ void pair_geometry_map_kernel(TrackSet ts, const int32_t *__restrict__ pairs, const int32_t *__restrict__ map, int n,
                                         int incr, int32_t *__restrict__ rows, int32_t *__restrict__ cols) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t m = map[k];
    if (m < 0) { rows[k] = 0; cols[k] = 0; return; }
    const int q = pairs[2 * m], r = pairs[2 * m + 1];
    rows[k] = (int)(ts.offsets[q + 1] - ts.offsets[q]) - incr;
    cols[k] = (int)(ts.offsets[r + 1] - ts.offsets[r]) - incr;
}
int launch_pair_geometry_map(const TrackSet &ts, const int32_t *pairs, const int32_t *map, int n, int incr, int32_t *rows,
                             int32_t *cols, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    pair_geometry_map_kernel<<<(n + 255) / 256, 256, 0, st>>>(ts, pairs, map, n, incr, rows, cols);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}
__global__ void scatter_scores_kernel(const float *__restrict__ src, const int32_t *__restrict__ map, int n, float *__restrict__ dst) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && map[k] >= 0) dst[map[k]] = src[k];
}
int launch_scatter_scores(const float *src, const int32_t *map, int n, float *dst, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    scatter_scores_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, map, n, dst);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

// Smith-Waterman over a CRP the pair pipeline emitted (BASELINE.json configs[1]: "Smith-Waterman on the same
// binary CRPs"): smith_waterman_constrained never reads the last row / column of its input
// (alignment_tools.py:36-39: B[i-1][j-1], i < M, j < N), so the DP matrix is (rows - 1) x (cols - 1) and the bit
// of column cols - 1 is cleared (the packed DP kernel relies on zero bits past its last column).
__global__ void __launch_bounds__(128) sw_trim_kernel(uint32_t *__restrict__ bits, int64_t slot_words, int wpr,
                                                      int32_t *__restrict__ rows, int32_t *__restrict__ cols) {
    const int k = blockIdx.x;
    const int R = rows[k], C = cols[k];
    if (C >= 1) {
        const int w = (C - 1) >> 5;
        const unsigned keep = ~(1u << ((C - 1) & 31));
        for (int i = threadIdx.x; i < R; i += blockDim.x) bits[(int64_t)k * slot_words + (int64_t)i * wpr + w] &= keep;
    }
    __syncthreads();
    if (threadIdx.x == 0) { rows[k] = R > 0 ? R - 1 : 0; cols[k] = C > 0 ? C - 1 : 0; }
}

__global__ void score_asymmetric_kernel(float *__restrict__ scores, const int32_t *__restrict__ cols, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) scores[k] = __fdiv_rn(__fsqrt_rn((float)cols[k]), scores[k]);
}
int launch_score_asymmetric(float *scores, const int32_t *cols, int n, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    score_asymmetric_kernel<<<(n + 255) / 256, 256, 0, st>>>(scores, cols, n);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

int launch_sw_trim(uint32_t *bits, int64_t slot_words, int words_per_row, int32_t *rows, int32_t *cols, int n,
                   cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    sw_trim_kernel<<<n, 128, 0, st>>>(bits, slot_words, words_per_row, rows, cols);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

int launch_pack_bytes(const uint8_t *mats, const int64_t *offsets, const int32_t *shapes, int n, int mode,
                      uint32_t *bits, int64_t slot_words, int words_per_row, int32_t *rows, int32_t *cols,
                      uint32_t *nonbinary_flag, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    const int max_rows = (int)(slot_words / words_per_row);
    dim3 grid((words_per_row + 7) / 8, max_rows, n);
    pack_bytes_kernel<<<grid, 256, 0, st>>>(mats, offsets, shapes, mode, bits, slot_words, words_per_row, rows, cols,
                                            nonbinary_flag);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

template <int MODE>
static int launch_packed_g(int G, const uint32_t *bits, int64_t slot_words, int wpr, const int32_t *rows,
                           const int32_t *cols, int n, float *scores, uint32_t *halo, int64_t halo_pitch,
                           cudaStream_t st) {
    const int blocks = (n + 3) / 4;
    switch (G) {
        case 1: dp_packed_kernel<1, MODE><<<blocks, 128, 0, st>>>(bits, slot_words, wpr, rows, cols, n, scores, halo, halo_pitch); break;
        case 2: dp_packed_kernel<2, MODE><<<blocks, 128, 0, st>>>(bits, slot_words, wpr, rows, cols, n, scores, halo, halo_pitch); break;
        default: dp_packed_kernel<3, MODE><<<blocks, 128, 0, st>>>(bits, slot_words, wpr, rows, cols, n, scores, halo, halo_pitch); break;
    }
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

// halo_scratch: at least n * 2 * halo_pitch * 16 bytes (float4 view for the scalar path)
int launch_dp_bits(const uint32_t *bits, int64_t slot_words, int words_per_row, const int32_t *rows,
                   const int32_t *cols, int n, int max_cols, int mode, float gamma_o, float gamma_e, float *scores,
                   uint32_t *halo_scratch, int64_t halo_pitch, cudaStream_t st, int64_t *launches) {
    if (n <= 0) return ACOSS_OK;
    if (launches) *launches += 1;
    const int max_rows = (int)(slot_words / words_per_row);
    const int lim = max_rows < max_cols ? max_rows : max_cols;      // scores are bounded by min(R, C)
    const bool packed_ok = (mode == ACOSS_ALIGN_QMAX) ? (gamma_o == 0.5f && gamma_e == 0.5f && lim <= 16383)
                                                      : (mode == ACOSS_ALIGN_SW && lim <= 3270);
    if (packed_ok) {
        const int dpcols = max_cols - 2;
        // 32-column groups per lane: least padded strip layout, ties to the widest strip
        int G = 1;
        long best_pad = -1;
        for (int cand = 1; cand <= 3; ++cand) {
            const long w = 1024L * cand, pad = (dpcols + w - 1) / w * w;
            if (best_pad < 0 || pad <= best_pad) { best_pad = pad; G = cand; }   // ties -> wider strips
        }
        if (mode == ACOSS_ALIGN_QMAX)
            return launch_packed_g<MODE_QMAX>(G, bits, slot_words, words_per_row, rows, cols, n, scores, halo_scratch,
                                              halo_pitch, st);
        return launch_packed_g<MODE_SW>(G, bits, slot_words, words_per_row, rows, cols, n, scores, halo_scratch,
                                        halo_pitch, st);
    }
    const int blocks = (n + 3) / 4;
    if (mode == ACOSS_ALIGN_DMAX && gamma_o == 0.5f && gamma_e == 0.5f && 2L * ((long)max_rows + max_cols) <= 32000L) {
        const int dpcols = max_cols - 3;
        if (dpcols <= 1024 || (dpcols > 2048 && dpcols <= 3072))   // least padded strip layout of 1024 / 2048 columns
            dp_dmax_packed_kernel<1><<<blocks, 128, 0, st>>>(bits, slot_words, words_per_row, rows, cols, n, scores,
                                                             (uint2 *)halo_scratch, halo_pitch);
        else
            dp_dmax_packed_kernel<2><<<blocks, 128, 0, st>>>(bits, slot_words, words_per_row, rows, cols, n, scores,
                                                             (uint2 *)halo_scratch, halo_pitch);
        CUDA_TRY(cudaGetLastError());
        return ACOSS_OK;
    }
    if (mode == ACOSS_ALIGN_QMAX)
        dp_scalar_kernel<MODE_QMAX><<<blocks, 128, 0, st>>>(bits, slot_words, words_per_row, rows, cols, n, gamma_o,
                                                            gamma_e, scores, (float4 *)halo_scratch, halo_pitch);
    else if (mode == ACOSS_ALIGN_SW)
        dp_scalar_kernel<MODE_SW><<<blocks, 128, 0, st>>>(bits, slot_words, words_per_row, rows, cols, n, gamma_o,
                                                          gamma_e, scores, (float4 *)halo_scratch, halo_pitch);
    else if (mode == ACOSS_ALIGN_DMAX)
        dp_dmax_kernel<1><<<blocks, 128, 0, st>>>(bits, slot_words, words_per_row, rows, cols, n, gamma_o, gamma_e,
                                                  scores, (float4 *)halo_scratch, halo_pitch);
    else if (mode == ACOSS_ALIGN_DMAX_PLAIN)
        dp_dmax_kernel<0><<<blocks, 128, 0, st>>>(bits, slot_words, words_per_row, rows, cols, n, gamma_o, gamma_e,
                                                  scores, (float4 *)halo_scratch, halo_pitch);
    else {
        acoss_set_error("alignment mode %d is not implemented", mode);
        return ACOSS_E_INVALID;
    }
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}
