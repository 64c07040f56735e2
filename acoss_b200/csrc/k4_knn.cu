// Row-only k-nearest-neighbour binarisation of float64 cross-similarity matrices.
//
// Replaces csm_to_binary(D, kappa) (/root/reference/acoss/algorithms/utils/cross_recurrence.py:137-161):
// exactly NN ones per row at the NN smallest entries (np.argpartition).  Ties at the NN-th value are
// broken by numpy's introselect in the reference; here the lowest column index wins (documented).
// One CTA per matrix row: the row is staged in shared memory as order-preserving 64-bit keys and
// the NN-th smallest key is found by an 8-pass, 8-bit radix select; a second ordered pass emits the
// bit-packed row.  nn < 0 means "all ones" (kappa == 0).
#include "common.cuh"

__device__ __forceinline__ unsigned long long dkey(double x) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);      // total order of finite doubles
}

__global__ void __launch_bounds__(256) knn_rows_kernel(const double *__restrict__ csms,
                                                       const int64_t *__restrict__ offsets,
                                                       const int32_t *__restrict__ shapes,
                                                       const int32_t *__restrict__ nn_a,
                                                       uint32_t *__restrict__ bits_dp, int64_t slot_words, int wpr,
                                                       uint32_t *__restrict__ bits_out,
                                                       const int64_t *__restrict__ out_offsets,
                                                       int32_t *__restrict__ rows, int32_t *__restrict__ cols) {
    extern __shared__ unsigned long long skeys[];
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned s_k, s_lt, s_warp_eq[8], s_run;
    const int k = blockIdx.y, i = blockIdx.x;
    const int M = shapes[2 * k], N = shapes[2 * k + 1];
    if (i == 0 && threadIdx.x == 0) { rows[k] = M - 1; cols[k] = N - 1; }   // SW never reads the last row/column
    if (i >= M) return;
    const int nn = nn_a[k];
    const double *row = csms + offsets[k] + (int64_t)i * N;
    for (int e = threadIdx.x; e < N; e += blockDim.x) skeys[e] = dkey(row[e]);
    unsigned long long vsel = 0ull;
    unsigned need_eq = 0;                                     // how many keys == vsel are taken
    bool all = nn < 0 || nn >= N, none = (nn == 0);
    if (!all && !none) {
        if (threadIdx.x == 0) { s_prefix = 0ull; s_k = (unsigned)(nn - 1); }
        unsigned long long mask = 0ull;
        for (int shift = 56; shift >= 0; shift -= 8) {
            hist[threadIdx.x] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            for (int e = threadIdx.x; e < N; e += blockDim.x) {
                const unsigned long long key = skeys[e];
                if ((key & mask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned cum = 0, kk = s_k;
                for (int b = 0; b < 256; ++b) {
                    const unsigned h = hist[b];
                    if (cum + h > kk) { s_prefix = prefix | ((unsigned long long)b << shift); s_k = kk - cum; break; }
                    cum += h;
                }
            }
            mask |= 255ull << shift;
            __syncthreads();
        }
        vsel = s_prefix;
        if (threadIdx.x == 0) s_lt = 0u;
        __syncthreads();
        unsigned lt = 0;
        for (int e = threadIdx.x; e < N; e += blockDim.x) lt += skeys[e] < vsel;
        atomicAdd(&s_lt, lt);
        __syncthreads();
        need_eq = (unsigned)nn - s_lt;
    } else {
        __syncthreads();
    }
    // ordered emission: columns in chunks of 256, ties (== vsel) taken in column order
    if (threadIdx.x == 0) s_run = 0u;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < wpr * 32; c0 += 256) {
        const int j = c0 + threadIdx.x;
        const bool in = j < N;
        const unsigned long long key = in ? skeys[j] : ~0ull;
        const bool eq = in && !all && !none && key == vsel;
        const unsigned eqm = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp_eq[warp] = __popc(eqm);
        __syncthreads();
        unsigned before = s_run;
        for (int w = 0; w < warp; ++w) before += s_warp_eq[w];
        before += __popc(eqm & ((1u << lane) - 1u));
        bool bit = in && (all || (!none && (key < vsel || (eq && before < need_eq))));
        const unsigned word = __ballot_sync(0xffffffffu, bit);
        const int w = (c0 >> 5) + warp;
        if (lane == 0 && w < wpr) {
            if (bits_out && w < (N + 31) / 32) bits_out[out_offsets[k] + (int64_t)i * ((N + 31) / 32) + w] = word;
            // DP copy: mask column N-1 and beyond (never read by smith_waterman_constrained)
            unsigned keep = 0xffffffffu;
            const int lim = N - 1 - w * 32;
            if (lim <= 0) keep = 0u; else if (lim < 32) keep = (1u << lim) - 1u;
            bits_dp[(int64_t)k * slot_words + (int64_t)i * wpr + w] = word & keep;
        }
        __syncthreads();
        if (threadIdx.x == 0) { unsigned t = 0; for (int w2 = 0; w2 < 8; ++w2) t += s_warp_eq[w2]; s_run += t; }
        __syncthreads();
    }
}

// Warp-per-row variant for rows of at most 32 * NT columns (the EarlyFusion block counts: a few hundred).
// Lane l keeps the keys of columns l, l + 32, ... in registers, so ballot(bit) over the warp IS the packed
// word of 32 consecutive columns.  The nn-th smallest key is found by a most-significant-bit-first binary
// radix select: per bit one REDUX (warp add) of the per-lane count of still-matching keys whose bit is 0;
// it stops as soon as a single key matches the prefix (distinct doubles: ~log2(N) bits past the common
// exponent).  With ties the full 64 bits run and the final rank inside the tie group says how many of the
// equal keys are taken — in column order, like the CTA kernel.  Same outputs as knn_rows_kernel.
template <int NT>
__global__ void __launch_bounds__(256) knn_rows_warp_kernel(const double *__restrict__ csms,
                                                            const int64_t *__restrict__ offsets,
                                                            const int32_t *__restrict__ shapes,
                                                            const int32_t *__restrict__ nn_a,
                                                            uint32_t *__restrict__ bits_dp, int64_t slot_words, int wpr,
                                                            uint32_t *__restrict__ bits_out,
                                                            const int64_t *__restrict__ out_offsets,
                                                            int32_t *__restrict__ rows, int32_t *__restrict__ cols) {
    const int k = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int M = shapes[2 * k], N = shapes[2 * k + 1];
    if (i == 0 && lane == 0) { rows[k] = M - 1; cols[k] = N - 1; }
    if (i >= M) return;
    const int nn = nn_a[k];
    const double *row = csms + offsets[k] + (int64_t)i * N;
    unsigned hi[NT], lo[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int j = lane + 32 * t;
        const unsigned long long key = j < N ? dkey(row[j]) : ~0ull;
        hi[t] = (unsigned)(key >> 32);
        lo[t] = (unsigned)key;
    }
    const bool all = nn < 0 || nn >= N, none = (nn == 0);
    unsigned vhi = 0u, vlo = 0u, need_eq = 0u;
    if (!all && !none) {
        unsigned alive = 0u;                                  // bit t: key t still matches the prefix
#pragma unroll
        for (int t = 0; t < NT; ++t) alive |= (lane + 32 * t < N) ? (1u << t) : 0u;
        unsigned kk = (unsigned)(nn - 1), alive_cnt = (unsigned)N;
        bool single = false;
        // bits on which all keys of the row agree (sign, exponent, leading mantissa bits of similar distances)
        // cannot split the set: only the differing bit positions are visited, most significant first
        unsigned ah = ~0u, oh = 0u, al = ~0u, ol = 0u;
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if ((alive >> t) & 1u) { ah &= hi[t]; oh |= hi[t]; al &= lo[t]; ol |= lo[t]; }
        const unsigned dh = __reduce_and_sync(0xffffffffu, ah) ^ __reduce_or_sync(0xffffffffu, oh);
        const unsigned dl = __reduce_and_sync(0xffffffffu, al) ^ __reduce_or_sync(0xffffffffu, ol);
        for (int half = 0; half < 2 && !single; ++half) {
            unsigned todo = half == 0 ? dh : dl;
            while (todo) {
                const int b = 31 - __clz(todo);
                todo &= ~(1u << b);
                unsigned zero = 0u;
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const unsigned w = half == 0 ? hi[t] : lo[t];
                    zero |= ((~w >> b) & 1u) << t;
                }
                zero &= alive;
                const unsigned total0 = __reduce_add_sync(0xffffffffu, (unsigned)__popc(zero));
                if (kk < total0) { alive = zero; alive_cnt = total0; }
                else { kk -= total0; alive &= ~zero; alive_cnt -= total0; }
                if (alive_cnt == 1u) { single = true; break; }
            }
        }
        // every surviving key is equal: take one of them as the threshold
        unsigned mh = 0u, ml = 0u;
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if ((alive >> t) & 1u) { mh = hi[t]; ml = lo[t]; }
        const unsigned owners = __ballot_sync(0xffffffffu, alive != 0u);
        const int src = __ffs(owners) - 1;
        vhi = __shfl_sync(0xffffffffu, mh, src);
        vlo = __shfl_sync(0xffffffffu, ml, src);
        need_eq = kk + 1u;                                    // rank inside the tie group (0 when single)
    }
    unsigned run = 0u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int wout = (N + 31) / 32;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        if (t >= wpr) break;
        const bool in = lane + 32 * t < N;
        const bool less = hi[t] < vhi || (hi[t] == vhi && lo[t] < vlo);
        const bool eq = in && !all && !none && hi[t] == vhi && lo[t] == vlo;
        const unsigned eqm = __ballot_sync(0xffffffffu, eq);
        const unsigned before = run + __popc(eqm & lt_mask);
        const bool bit = in && (all || (!none && (less || (eq && before < need_eq))));
        const unsigned word = __ballot_sync(0xffffffffu, bit);
        run += __popc(eqm);
        if (lane == 0) {
            if (bits_out && t < wout) bits_out[out_offsets[k] + (int64_t)i * wout + t] = word;
            unsigned keep = 0xffffffffu;                      // DP copy: column N-1 and beyond are never read
            const int lim = N - 1 - t * 32;
            if (lim <= 0) keep = 0u; else if (lim < 32) keep = (1u << lim) - 1u;
            bits_dp[(int64_t)k * slot_words + (int64_t)i * wpr + t] = word & keep;
        }
    }
    // pad words beyond the register-resident columns (wpr = ceil(max_cols / 32) + 1 <= NT + 1)
    if (lane == 0)
        for (int t = NT; t < wpr; ++t) bits_dp[(int64_t)k * slot_words + (int64_t)i * wpr + t] = 0u;
}

int launch_knn_rows(const double *csms, const int64_t *offsets, const int32_t *shapes, const int32_t *nn, int n,
                    int max_rows, int max_cols, uint32_t *bits_dp, int64_t slot_words, int wpr, uint32_t *bits_out,
                    const int64_t *out_offsets, int32_t *rows, int32_t *cols, cudaStream_t st) {
    if (n <= 0) return ACOSS_OK;
    if (max_cols <= 1024) {                                   // rows fit the registers of one warp
        dim3 grid((unsigned)((max_rows + 7) / 8), (unsigned)n);
        if (max_cols <= 256)
            knn_rows_warp_kernel<8><<<grid, 256, 0, st>>>(csms, offsets, shapes, nn, bits_dp, slot_words, wpr, bits_out, out_offsets, rows, cols);
        else if (max_cols <= 512)
            knn_rows_warp_kernel<16><<<grid, 256, 0, st>>>(csms, offsets, shapes, nn, bits_dp, slot_words, wpr, bits_out, out_offsets, rows, cols);
        else
            knn_rows_warp_kernel<32><<<grid, 256, 0, st>>>(csms, offsets, shapes, nn, bits_dp, slot_words, wpr, bits_out, out_offsets, rows, cols);
        CUDA_TRY(cudaGetLastError());
        return ACOSS_OK;
    }
    const size_t smem = (size_t)max_cols * 8;
    if (smem > 200 * 1024) { acoss_set_error("knn: rows longer than %d columns are not supported", 200 * 1024 / 8); return ACOSS_E_INVALID; }
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(knn_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_rows_kernel<<<dim3(max_rows, n), 256, smem, st>>>(csms, offsets, shapes, nn, bits_dp, slot_words, wpr, bits_out,
                                                         out_offsets, rows, cols);
    CUDA_TRY(cudaGetLastError());
    return ACOSS_OK;
}
