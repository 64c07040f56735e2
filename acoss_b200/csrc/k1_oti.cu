// K1 — per-track global chroma and batched optimal-transposition index (OTI).
//
// Replaces essentia optimalTranspositionIndex / globalAverageChroma inside ChromaCrossSimilarity
// (called from /root/reference/acoss/algorithms/rqa_serra09.py:66):
//   g[b]  = (sequential float32 sum over frames of x[f][b]) / max_b(...)      (skip if max == 0)
//   oti   = first argmax_{s=0..noti} dot(g_q, rotR(g_r, s)),  rotR(x,s)[b] = x[(b-s) mod 12]
//   dot   = float32 products accumulated sequentially in float64, narrowed to float32 (F3)
// Layout: one 16-lane sub-warp per track / per pair; lanes 0..11 own the 12 chroma bins (global
// chroma) or lanes 0..15 own circular shifts (OTI), and the argmax is a shuffle reduction.
#include "common.cuh"

__global__ void __launch_bounds__(256) global_chroma_kernel(const float *__restrict__ frames,
                                                            const int64_t *__restrict__ offsets,
                                                            int n_tracks, float *__restrict__ gchroma) {
    const int sub = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;   // 16-lane group = one track
    const int lane = threadIdx.x & 15;
    const bool live = sub < n_tracks;
    float acc = 0.f;
    if (live && lane < NBINS) {
        const int64_t f0 = offsets[sub], f1 = offsets[sub + 1];
        const float *p = frames + f0 * NBINS + lane;
        for (int64_t f = f0; f < f1; ++f, p += NBINS) acc = __fadd_rn(acc, *p);   // strictly sequential
    }
    float mx = (lane < NBINS) ? acc : -INFINITY;
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o, 16));
    if (live && lane < NBINS) gchroma[sub * NBINS + lane] = (mx != 0.f) ? __fdiv_rn(acc, mx) : acc;
}

__global__ void __launch_bounds__(256) oti_kernel(TrackSet ts, const int32_t *__restrict__ pairs,
                                                  int64_t n_pairs, int noti, int apply,
                                                  int32_t *__restrict__ oti_out) {
    const int64_t sub = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;   // one pair per 16 lanes
    const int lane = threadIdx.x & 15;
    const bool live = sub < n_pairs;
    float gq[NBINS], gr[NBINS];
    if (live) {
        const int q = pairs[2 * sub], r = pairs[2 * sub + 1];
#pragma unroll
        for (int b = 0; b < NBINS; ++b) {
            gq[b] = ts.gchroma[q * NBINS + b];
            gr[b] = ts.gchroma[r * NBINS + b];
        }
    } else {
#pragma unroll
        for (int b = 0; b < NBINS; ++b) gq[b] = gr[b] = 0.f;
    }
    // each lane scores shifts s = lane, lane+16, ... <= noti ; keeps its first maximum
    float best = -INFINITY;
    int best_s = 0x7fffffff;
    for (int s = lane; s <= noti; s += 16) {
        const int sm = s % NBINS;
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < NBINS; ++b) {
            // gr[(b - sm) mod 12] without dynamic register indexing
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < NBINS; ++k) v = (rot_src(b, sm) == k) ? gr[k] : v;
            acc = acc_f32prod(acc, gq[b], v);
        }
        const float val = (float)acc;
        if (val > best) { best = val; best_s = s; }
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o, 16);
        const int os = __shfl_xor_sync(0xffffffffu, best_s, o, 16);
        if (ob > best || (ob == best && os < best_s)) { best = ob; best_s = os; }
    }
    if (live && lane == 0) oti_out[sub] = apply ? (best_s == 0x7fffffff ? 0 : best_s) : 0;
}

// max squared frame norm and max feature value (as float bits: non-negative floats order like ints), min feature value
__global__ void __launch_bounds__(256) frame_stats_kernel(const float *__restrict__ frames, int64_t total_frames,
                                                          float *__restrict__ stats2) {
    float mx = 0.f, mn = 0.f, mc = 0.f;
    bool bad = false;                                         // NaN / Inf features: fmaxf / fminf would drop them silently
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total_frames; f += (int64_t)gridDim.x * blockDim.x) {
        float n2 = 0.f;
#pragma unroll
        for (int b = 0; b < NBINS; ++b) {
            const float v = frames[f * NBINS + b];
            n2 = fmaf(v, v, n2);
            mn = fminf(mn, v);
            mc = fmaxf(mc, v);
        }
        bad |= !(n2 <= 3.0e38f);
        mx = fmaxf(mx, n2);
    }
    if (__any_sync(0xffffffffu, bad)) mx = __int_as_float(0x7f800000);   // +inf => no fixed point => exact path (flags the NaN)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mc = fmaxf(mc, __shfl_xor_sync(0xffffffffu, mc, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax((int *)&stats2[0], __float_as_int(mx));
        if (mn < 0.f) atomicExch((int *)&stats2[1], __float_as_int(-1.f));
        atomicMax((int *)&stats2[2], __float_as_int(mc));
    }
}

int launch_frame_stats(const float *frames, int64_t total_frames, float *stats2, cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(stats2, 0, 12, st));
    if (total_frames <= 0) return ACOSS_OK;
    frame_stats_kernel<<<148 * 4, 256, 0, st>>>(frames, total_frames, stats2);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

int launch_global_chroma(const float *frames, const int64_t *offsets, int n_tracks, float *gchroma,
                         cudaStream_t st) {
    if (n_tracks <= 0) return ACOSS_OK;
    const int threads = 256, per = threads / 16;
    global_chroma_kernel<<<(n_tracks + per - 1) / per, threads, 0, st>>>(frames, offsets, n_tracks, gchroma);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}

int launch_oti(const TrackSet &ts, const int32_t *pairs, int64_t n_pairs, int noti, int apply,
               int32_t *oti_out, cudaStream_t st) {
    if (n_pairs <= 0) return ACOSS_OK;
    const int threads = 256, per = threads / 16;
    const int64_t blocks = (n_pairs + per - 1) / per;
    oti_kernel<<<(unsigned)blocks, threads, 0, st>>>(ts, pairs, n_pairs, noti, apply, oti_out);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}
