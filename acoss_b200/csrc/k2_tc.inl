// Tensor-core sweeps of the fast CRP path (included by k2_fast.cu, inside its anonymous namespace).
//
// The item of cell (window r of the streamed track, window j of the owned track) is
//     z = xn[r] + yn[j] - 2 <X_r, Y_j>,      X_r = frames r .. r+8 stacked (108 values),
// and the 108-term inner product is a GEMM the 5th-generation tensor cores do exactly in integers:
//   * prep quantises every feature to 24 bits, x_q = rint(x * 2^q_exp) < 2^24, and stores the three byte limbs
//     (h, l1, l2) as planes of 16-byte frames (12 bins + 4 zero bytes);
//   * <X_q, Y_q> = 2^32 hh + 2^24 (h.l1 + l1.h) + 2^16 (l1.l1 + h.l2 + l2.h) + (dropped: 2^8 (l1.l2 + l2.l1) + l2.l2):
//     three int32 accumulators in TMEM, filled by tcgen05.mma.kind::i8 (u8 x u8 -> s32, M = 128 owned windows,
//     N = 64 streamed windows, K = 32 bytes = two 16-byte frames);
//   * a stacked operand row is never materialised: with the no-swizzle K-major layout a core matrix is 8 rows x 16
//     bytes, row r / tap t is frame r + t, so core matrix (rows r0 .. r0+7, tap t) IS the 128 contiguous bytes at
//     frame r0 + t of the plane: leading byte offset 16 (next tap), stride byte offset 128 (next 8 rows).  One
//     MMA covers two taps; the ninth tap pairs with a block of zeros on the streamed side.  30 MMAs per block;
//   * consumer threads read their owned window's 64 items from TMEM (tcgen05.ld 32x32b: TMEM lane = owned window,
//     column = streamed window) and bin / classify them: no dot products, no sliding sums, no shuffles on the
//     CUDA cores.
// T = (acc0 << s0) + (acc1 >> s1) + (acc2 >> s2) is 2 <X, Y> in the fixed-point unit of the path (shifts from q_exp and
// fx_exp); the sparse level evaluates the same integer formula with dp4a, so every kernel sees identical items.
// Error budget against the exact item: quantisation 216 x_max 2^-(q_exp+1) * 2, dropped limb products
// 2 * 108 * 255^2 * 2^(9 - 2 q_exp), two floors: 36 units at HPCP scale, inside EPS.

struct TcShift { int s0, s1, s2; };

constexpr int TC_AF = 144;                          // owned frames held per plane (128 windows + 9 taps, rounded)
constexpr int TC_ALOAD = 137;                       // owned frames loaded per plane
constexpr int TC_BF = 72;                           // streamed frames per plane and stage (64 windows + 8)
constexpr int TC_CONS = 256;                        // consumer threads: warp w reads TMEM lanes 32 (w & 3) .., block half w >> 2
constexpr int TC_THREADS = TC_CONS + 32;            // + one producer warp (TMA, MMA issue, TMEM allocation)
constexpr int TC_TMEM_COLS = 256;                   // three accumulators of 64 columns (allocation: power of two)

struct alignas(128) TcSmem {
    uint8_t a[3][TC_AF][16];                        // owned planes h, l1, l2
    uint8_t b[2][3][TC_BF][16];                     // streamed planes, two stages
    uint8_t zero[TC_N][16];
    unsigned long long afull, bfull[2], bfree[2], acc_full, acc_free;
    uint32_t tmem_base;
};

// K-major, no swizzle: ((8, m), 2) : ((16 B, SBO), LBO); version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)(128u >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor: s32 accumulators, u8 x u8, both K-major, M = 128, N = 64
constexpr uint32_t TC_IDESC = (2u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// the 30 MMAs of one block: acc0 = h.h, acc1 = h.l1 + l1.h, acc2 = l1.l1 + h.l2 + l2.h  (owned plane, streamed plane)
__device__ __forceinline__ void tc_issue_block(uint32_t a0, uint32_t b0, uint32_t zero, uint32_t tmem) {
    constexpr uint32_t PA = TC_AF * 16, PB = TC_BF * 16;
    auto prod = [&](uint32_t pa, uint32_t pb, uint32_t acc, bool first) {
        const uint32_t sa = a0 + pa * PA, sb = b0 + pb * PB;
#pragma unroll
        for (int t = 0; t < 8; t += 2) tc_mma(tmem + acc * TC_N, tc_desc(sa + 16 * t, 16), tc_desc(sb + 16 * t, 16), !(first && t == 0));
        tc_mma(tmem + acc * TC_N, tc_desc(sa + 16 * 8, 16), tc_desc(sb + 16 * 8, zero - (sb + 16 * 8)), 1u);   // tap 8 x (tap 8, zeros)
    };
    prod(0, 0, 0, true);
    prod(0, 1, 1, true);
    prod(1, 0, 1, false);
    prod(1, 1, 2, true);
    prod(0, 2, 2, false);
    prod(2, 0, 2, false);
}

// Producer warp (one elected lane): owned planes once, then per block the streamed planes (double buffered) and the MMAs.
__device__ __forceinline__ void tc_producer(TcSmem *s, const uint8_t *__restrict__ own, const uint8_t *__restrict__ str,
                                            size_t plane_bytes, int own_first, int nblocks, uint32_t tmem) {
    auto load_b = [&](int b) {
        const int st = b & 1;
        mbar_expect_tx(&s->bfull[st], 3 * TC_BF * 16);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl)
            bulk_g2s(&s->b[st][pl][0][0], str + pl * plane_bytes + (size_t)b * TC_N * 16, TC_BF * 16, &s->bfull[st]);
    };
    mbar_expect_tx(&s->afull, 3 * TC_ALOAD * 16);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) bulk_g2s(&s->a[pl][0][0], own + pl * plane_bytes + (size_t)own_first * 16, TC_ALOAD * 16, &s->afull);
    load_b(0);
    if (nblocks > 1) load_b(1);
    mbar_wait(&s->afull, 0);
    const uint32_t a0 = smem_u32(&s->a[0][0][0]), zero = smem_u32(&s->zero[0][0]);
    for (int b = 0; b < nblocks; ++b) {
        const int st = b & 1;
        mbar_wait(&s->bfull[st], (uint32_t)((b >> 1) & 1));
        if (b > 0) mbar_wait(&s->acc_free, (uint32_t)((b - 1) & 1));      // consumers hold block b - 1 in registers
        tc_fence_after();
        tc_issue_block(a0, smem_u32(&s->b[st][0][0][0]), zero, tmem);
        tc_commit(&s->acc_full);
        tc_commit(&s->bfree[st]);
        if (b + 2 < nblocks) {
            mbar_wait(&s->bfree[st], (uint32_t)((b >> 1) & 1));           // the MMAs of block b have read stage st
            load_b(b + 2);
        }
    }
}

// Common CTA prologue / epilogue: barriers, zero block, TMEM allocation by the producer warp.
__device__ __forceinline__ uint32_t tc_begin(TcSmem *s) {
    const int tid = threadIdx.x;
    for (int i = tid; i < TC_N * 16 / 4; i += TC_THREADS) reinterpret_cast<uint32_t *>(&s->zero[0][0])[i] = 0u;
    if (tid == 0) {
        mbar_init(&s->afull, 1);
        mbar_init(&s->bfull[0], 1); mbar_init(&s->bfull[1], 1);
        mbar_init(&s->bfree[0], 1); mbar_init(&s->bfree[1], 1);
        mbar_init(&s->acc_full, 1);
        mbar_init(&s->acc_free, TC_CONS);
        mbar_fence_init();
    }
    if (tid >= TC_CONS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s->tmem_base)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // the zero block is read by the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return s->tmem_base;
}
__device__ __forceinline__ void tc_end(uint32_t tmem) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x >= TC_CONS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
}

// Consumer side of one block: fn(chunk_first_window, acc0[16], acc1[16], acc2[16]) for the two 16-window chunks of this warp's
// half.  The accumulators are released (acc_free) as soon as the last chunk sits in registers.
template <typename Fn>
__device__ __forceinline__ void tc_consume_block(TcSmem *s, uint32_t tmem, int b, int warp, Fn &&fn) {
    mbar_wait(&s->acc_full, (uint32_t)(b & 1));
    tc_fence_after();
    const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        int v0[16], v1[16], v2[16];
        tc_ld16(tbase + ch * 16, v0);
        tc_ld16(tbase + TC_N + ch * 16, v1);
        tc_ld16(tbase + 2 * TC_N + ch * 16, v2);
        tc_ld_wait();
        if (ch == 1) {
            tc_fence_before();
            mbar_arrive(&s->acc_free);
        }
        fn(b * TC_N + (warp >> 2) * 32 + ch * 16, v0, v1, v2);
    }
}

// ------------------------------------------------------------------------------------------------
// histogram sweep on the tensor cores.  ORIENT as in fast_hist_kernel: 0 = owned reference columns, streamed query
// windows; 1 = owned query rows, streamed reference windows.  One CTA = 128 owned lines of one pair.
// ------------------------------------------------------------------------------------------------
template <int ORIENT>
__global__ void __launch_bounds__(TC_THREADS, 2) tc_hist_kernel(TrackSet ts, const int32_t *__restrict__ pairs, int64_t first, int n,
                                                                FastLayout L, char *__restrict__ scratch, int strips_max, TcShift sh3,
                                                                uint32_t *__restrict__ status, uint32_t *__restrict__ dbg, int min_live,
                                                                int final_level, uint32_t *__restrict__ glive, uint32_t gcap,
                                                                int32_t *__restrict__ dbgz) {
    extern __shared__ __align__(128) unsigned char tc_raw[];
    TcSmem *s = reinterpret_cast<TcSmem *>(tc_raw);
    uint32_t *hist = reinterpret_cast<uint32_t *>(tc_raw + sizeof(TcSmem));   // [NBIN + 2][128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = blockIdx.x / strips_max, strip = blockIdx.x - slot * strips_max;
    if (slot >= n) return;
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    if (h->quirk[ORIENT == 0 ? 1 : 0]) return;               // threshold forced to 0: nothing to select
    const int64_t k = first + slot;
    const int nY = (ORIENT == 0) ? h->nr : h->nq, nX = (ORIENT == 0) ? h->nq : h->nr;
    const int My = nY - M9, Mxs = nX - M9;                    // owned / streamed windows
    const int cb = strip * 128;
    if (cb >= My) return;
    const int32_t *xn = slot_ptr<int32_t>(scratch, L, slot, ORIENT == 0 ? L.off_aai : L.off_bbi);
    const int32_t *yn = slot_ptr<int32_t>(scratch, L, slot, ORIENT == 0 ? L.off_bbi : L.off_aai);
    const int line0 = (ORIENT == 0) ? L.max_rows : 0;
    int32_t *lo_a = slot_ptr<int32_t>(scratch, L, slot, L.off_lo) + line0;
    int32_t *w_a = slot_ptr<int32_t>(scratch, L, slot, L.off_w) + line0;
    int32_t *cb_a = slot_ptr<int32_t>(scratch, L, slot, L.off_cb) + line0;
    int32_t *sh_a = slot_ptr<int32_t>(scratch, L, slot, L.off_sh) + line0;
    const int side = (ORIENT == 0) ? 1 : 0;
    const bool cons = tid < TC_CONS;
    const int tl = (warp & 3) * 32 + lane;                    // owned line within the CTA (consumers)
    const int j = cb + tl;
    bool valid = cons && j < My;
    const int shv = valid ? sh_a[j] : -1;
    if (shv < 0) valid = false;
    const int shf = valid ? shv : 0;
    // bin = ((z - lo) >> sh) + 1 clamped to [0, NBIN + 1]; idle lines land in the overflow bin
    const int ynrel = valid ? yn[j] - lo_a[j] + (1 << shv) : 0x40000000;
    const bool scan = valid && warp < 4;                     // one thread per line scans (both halves count into one histogram)
    const int n_live_cta = __syncthreads_count(scan);
    if (n_live_cta == 0) return;
    auto to_sparse = [&](int line) {
        const unsigned pos = atomicAdd(glive, 1u);
        if (pos < gcap) glive[1 + pos] = ((uint32_t)slot << 16) | ((uint32_t)side << 15) | (uint32_t)line;
        else atomicOr(&status[k], PAIR_ST_FALLBACK | 8u);
    };
    if (n_live_cta < min_live) {
        if (scan) to_sparse(j);
        return;
    }
    for (int i = tid; i < (NBIN + 2) * 128; i += TC_THREADS) hist[i] = 0u;
    const uint32_t tmem = tc_begin(s);
    const int nblocks = (Mxs + TC_N - 1) / TC_N;
    if (!cons) {
        if (lane == 0) {
            const uint8_t *qpl = slot_ptr<uint8_t>(scratch, L, slot, L.off_qpl), *rpl = slot_ptr<uint8_t>(scratch, L, slot, L.off_rpl);
            tc_producer(s, ORIENT == 0 ? rpl : qpl, ORIENT == 0 ? qpl : rpl, (size_t)L.plane_frames * 16, cb, nblocks, tmem);
        }
    } else {
        uint32_t *hp = hist + tl;
        for (int b = 0; b < nblocks; ++b) {
            tc_consume_block(s, tmem, b, warp, [&](int r0, const int (&v0)[16], const int (&v1)[16], const int (&v2)[16]) {
                const int4 *xp = reinterpret_cast<const int4 *>(xn + r0);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int4 xv = __ldg(xp + g);
                    const int xb[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = 4 * g + u;
                        const int T = (v0[i] << sh3.s0) + (v1[i] >> sh3.s1) + (v2[i] >> sh3.s2);
                        const int zr = xb[u] + ynrel - T;
                        const int idx = __vimin_s32_relu(zr >> shf, NBIN + 1);
                        atomicAdd(&hp[idx * 128], 1u);
                        if (dbgz && slot == 0 && strip == 0 && r0 + i < 64) dbgz[tl * 64 + r0 + i] = valid ? xb[u] + yn[j] - T : -1;
                    }
                }
            });
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_CONS) : "memory");
        int n_live = scan ? 1 : 0, n_miss = 0, n_left = 0;
        if (scan) {
            const int fk = h->fk[side], ck = h->ck[side];
            const Bracket br = split_bracket<128>(hp, 0, 0xffffffffu, fk, ck, lo_a[j], shf, h->lo1, h->hi1, bracket_target(nX - M9));
            if (br.bad) {
                atomicOr(&status[k], PAIR_ST_FALLBACK | 4u);
                sh_a[j] = -1;
            } else {
                lo_a[j] = br.lo; w_a[j] = br.w; cb_a[j] = br.below; sh_a[j] = br.done ? -1 : br.sh;
                n_miss = br.miss ? 1 : 0; n_left = br.done ? 0 : 1;
                if (ORIENT == 1) {
                    int4 *rowpack = slot_ptr<int4>(scratch, L, slot, L.off_rowpack);
                    rowpack[j] = make_int4(yn[j], br.lo - 2 * EPS, br.w + 4 * EPS, 0);
                }
                if (!br.done && final_level) to_sparse(j);
            }
        }
        if (warp < 4) {
            n_live = __reduce_add_sync(0xffffffffu, n_live);
            n_miss = __reduce_add_sync(0xffffffffu, n_miss);
            n_left = __reduce_add_sync(0xffffffffu, n_left);
            if (lane == 0) {
                if (warp == 0) atomicAdd(&dbg[0], 1u);
                atomicAdd(&dbg[1], (unsigned)n_live);
                if (n_miss) atomicAdd(&dbg[2], (unsigned)n_miss);
                if (n_left) atomicAdd(&dbg[3], (unsigned)n_left);
            }
        }
    }
    tc_end(tmem);
}

// ------------------------------------------------------------------------------------------------
// emit sweep on the tensor cores (owned = reference columns, streamed = query rows).  One CTA = 128 CRP columns = 4
// whole CRP words per row.  Per cell: the item from TMEM, the classification of fast_emit_kernel ("certainly 1" /
// "uncertain": inside a row or column bracket widened by 2 EPS, or near zero), one ballot per row = the CRP word of
// the warp's 32 columns; uncertain cells are staged per lane and compacted to the pair's pool once per 32 rows.
// ------------------------------------------------------------------------------------------------
constexpr int TC_STAGE = 32;                        // staged records per lane and half block (every cell fits)

__global__ void __launch_bounds__(TC_THREADS, 2) tc_emit_kernel(TrackSet ts, const int32_t *__restrict__ pairs, int64_t first, int n,
                                                                FastLayout L, char *__restrict__ scratch, int groups, TcShift sh3,
                                                                uint32_t *__restrict__ crp_all, int words, int64_t crp_words) {
    extern __shared__ __align__(128) unsigned char tc_raw[];
    TcSmem *s = reinterpret_cast<TcSmem *>(tc_raw);
    uint2 *stage_all = reinterpret_cast<uint2 *>(tc_raw + sizeof(TcSmem));    // [8 warps][TC_STAGE][32 lanes]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = blockIdx.x / groups, grp = blockIdx.x - slot * groups;
    if (slot >= n) return;
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    const int nY = h->nr, nX = h->nq, My = nY - M9, Mx = nX - M9;
    uint32_t *crp = crp_all + (int64_t)slot * crp_words;
    const int cb = grp * 128, w_first = grp * 4;
    if (cb >= My) {
        // no columns: the CTA's words of the pair's rows are zero (K3 reads the whole row pitch)
        const int nw = min(words, w_first + 4) - w_first;
        if (nw <= 0) return;
        for (int e = tid; e < Mx * nw; e += TC_THREADS) {
            const int i = e / nw, w = w_first + (e - i * nw);
            crp[(int64_t)i * words + w] = 0u;
        }
        return;
    }
    const uint32_t tmem = tc_begin(s);
    const int nblocks = (Mx + TC_N - 1) / TC_N;
    if (tid >= TC_CONS) {
        if (lane == 0)
            tc_producer(s, slot_ptr<uint8_t>(scratch, L, slot, L.off_rpl), slot_ptr<uint8_t>(scratch, L, slot, L.off_qpl),
                        (size_t)L.plane_frames * 16, cb, nblocks, tmem);
    } else {
        const int32_t *yn = slot_ptr<int32_t>(scratch, L, slot, L.off_bbi);
        const int4 *rowpack = slot_ptr<int4>(scratch, L, slot, L.off_rowpack);
        const int32_t *lo_c = slot_ptr<int32_t>(scratch, L, slot, L.off_lo) + L.max_rows;
        const int32_t *w_c = slot_ptr<int32_t>(scratch, L, slot, L.off_w) + L.max_rows;
        uint2 *pool = slot_ptr<uint2>(scratch, L, slot, L.off_pool);
        uint32_t *pool_ctr = slot_ptr<uint32_t>(scratch, L, slot, L.off_pcnt);
        const unsigned pool_cap = (unsigned)L.pool_cap;
        const int qd = warp & 3, hf = warp >> 2;
        const int j = cb + qd * 32 + lane;                    // CRP column of this thread
        const bool valid = j < My;
        const int ynv = valid ? yn[j] : TC_HUGE;              // invalid: item huge => never in, never uncertain
        const int ycl = valid ? yn[j] - (lo_c[j] - 2 * EPS) : TC_HUGE;
        const unsigned cw1 = valid ? (unsigned)(w_c[j] + 4 * EPS - 1) : 0u;
        const unsigned jrec = (unsigned)j << 14;
        const int wcol = w_first + qd;                        // CRP word of this warp's 32 columns
        uint2 *stage = stage_all + (size_t)warp * TC_STAGE * 32 + lane;
        for (int b = 0; b < nblocks; ++b) {
            unsigned word = 0u, ns = 0u;
            tc_consume_block(s, tmem, b, warp, [&](int r0, const int (&v0)[16], const int (&v1)[16], const int (&v2)[16]) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int4 rp = __ldg(rowpack + r0 + i);  // {aa_fix, rowLo - 2 EPS, rowW + 4 EPS, -}
                    const int T = (v0[i] << sh3.s0) + (v1[i] >> sh3.s1) + (v2[i] >> sh3.s2);
                    const int ar = rp.x - rp.y + ynv - T;
                    const int ac = rp.x + ycl - T;
                    const bool unc = ((unsigned)ar <= (unsigned)(rp.z - 1)) || ((unsigned)ac <= cw1) || (ar < 2 * EPS - rp.y);
                    const unsigned bal = __ballot_sync(0xffffffffu, (ar & ac) < 0);
                    if (lane == ((r0 + i) & 31)) word = bal;
                    if (unc) {
                        stage[ns * 32] = make_uint2((unsigned)(r0 + i) | jrec, (unsigned)(ar + rp.y));
                        ++ns;
                    }
                }
            });
            // lane l holds the word of row b * 64 + hf * 32 + l
            const int row = b * TC_N + hf * 32 + lane;
            if (row < Mx && wcol < words) crp[(int64_t)row * words + wcol] = word;
            if (__any_sync(0xffffffffu, ns != 0u)) {
                const unsigned cnt = ns;
                unsigned incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                unsigned base = 0u;
                if (lane == 31) base = atomicAdd(pool_ctr, incl);
                base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
#pragma unroll 1
                for (unsigned e = 0; e < cnt; ++e)
                    if (base + e < pool_cap) pool[base + e] = stage[e * 32];
            }
        }
    }
    tc_end(tmem);
}

// ------------------------------------------------------------------------------------------------
// sparse refinement with the integer items of the tensor sweeps: the three limb products as dp4a chains over the byte
// planes, combined by the same shifts, so a line's histogram counts exactly the items the sweeps see.
// Structure as fast_sparse_kernel: one lane = one crowded line, the 9 owned frames in registers.
// ------------------------------------------------------------------------------------------------
struct Limbs { uint32_t w[3][3]; };                 // [plane h, l1, l2][12 bytes]

__device__ __forceinline__ void tc_load_limbs(const uint4 *__restrict__ planes, int plane_frames, int f, Limbs &o) {
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const uint4 v = __ldg(planes + (size_t)pl * plane_frames + f);
        o.w[pl][0] = v.x; o.w[pl][1] = v.y; o.w[pl][2] = v.z;
    }
}
__device__ __forceinline__ uint32_t dp12(const uint32_t (&a)[3], const uint32_t (&b)[3], uint32_t c) {
    return __dp4a(a[2], b[2], __dp4a(a[1], b[1], __dp4a(a[0], b[0], c)));
}
template <int U>
__device__ __forceinline__ void tc_sparse_step(const Limbs (&y)[M9], uint32_t (&a0)[M9], uint32_t (&a1)[M9], uint32_t (&a2)[M9], const Limbs &x) {
#pragma unroll
    for (int t = 0; t < M9; ++t) {
        const int sl = ((U - t) % M9 + M9) % M9;              // row a - t lives in slot (a - t) % 9
        const uint32_t e0 = dp12(y[t].w[0], x.w[0], 0u);
        const uint32_t e1 = dp12(y[t].w[1], x.w[0], dp12(y[t].w[0], x.w[1], 0u));
        const uint32_t e2 = dp12(y[t].w[2], x.w[0], dp12(y[t].w[0], x.w[2], dp12(y[t].w[1], x.w[1], 0u)));
        if (t == 0) { a0[sl] = e0; a1[sl] = e1; a2[sl] = e2; }
        else { a0[sl] += e0; a1[sl] += e1; a2[sl] += e2; }
    }
}

__global__ void __launch_bounds__(32 * WPC, 2) tc_sparse_kernel(TrackSet ts, const int32_t *__restrict__ pairs, int64_t first, int n,
                                                                FastLayout L, char *__restrict__ scratch, TcShift sh3,
                                                                uint32_t *__restrict__ status, uint32_t *__restrict__ dbg,
                                                                const uint32_t *__restrict__ glive, uint32_t gcap) {
    __shared__ uint32_t s_sp[WPC][NBIN + 2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cnt = min(glive[0], gcap);
    const uint32_t e0 = ((uint32_t)blockIdx.x * WPC + warp) * 32u;
    if (e0 >= cnt) return;
    bool livel = e0 + lane < cnt;
    const uint32_t ent = glive[1 + (livel ? e0 + lane : e0)];
    const int slot = (int)(ent >> 16), side = (int)((ent >> 15) & 1u), j = (int)(ent & 0x7fffu);   // side 0: rows (owned = query)
    const int64_t k = first + slot;
    const PairHdr *h = slot_ptr<PairHdr>(scratch, L, slot, L.off_hdr);
    const int nX = side ? h->nq : h->nr;
    const uint4 *qpl = slot_ptr<uint4>(scratch, L, slot, L.off_qpl), *rpl = slot_ptr<uint4>(scratch, L, slot, L.off_rpl);
    const uint4 *Y = side ? rpl : qpl, *X = side ? qpl : rpl;
    const int32_t *xn = slot_ptr<int32_t>(scratch, L, slot, side ? L.off_aai : L.off_bbi);
    const int32_t *yn = slot_ptr<int32_t>(scratch, L, slot, side ? L.off_bbi : L.off_aai);
    const int line0 = side ? L.max_rows : 0;
    int32_t *lo_a = slot_ptr<int32_t>(scratch, L, slot, L.off_lo) + line0;
    int32_t *w_a = slot_ptr<int32_t>(scratch, L, slot, L.off_w) + line0;
    int32_t *cb_a = slot_ptr<int32_t>(scratch, L, slot, L.off_cb) + line0;
    int32_t *sh_a = slot_ptr<int32_t>(scratch, L, slot, L.off_sh) + line0;
    Limbs y[M9];
#pragma unroll
    for (int t = 0; t < M9; ++t) tc_load_limbs(Y, L.plane_frames, j + t, y[t]);   // frames j .. j+8 exist (j is a window)
    const int ynj = yn[j];
    const int fk = h->fk[side], ck = h->ck[side], rlo = h->lo1, rhi = h->hi1;
    uint32_t *hist = &s_sp[warp][0][lane];
    int lo = lo_a[j], sh = sh_a[j];
    if (sh < 0) livel = false;
    const int nrows = nX - 1;                                 // this lane's streamed frames 0 .. nX-2
    const int nrows_w = __reduce_max_sync(0xffffffffu, nrows);
    int n_swept = 0;
    for (int lvl = 0; lvl < SPARSE_LEVELS && __any_sync(0xffffffffu, livel); ++lvl) {
        ++n_swept;
#pragma unroll 1
        for (int b = 0; b < NBIN + 2; ++b) hist[b * 32] = 0u;
        const int shl = max(sh, 0);
        const int yrel = livel ? ynj - lo + (1 << shl) : 0x40000000;   // idle lanes land in the overflow bin
        int fx = 0;
        Limbs xc;
        tc_load_limbs(X, L.plane_frames, fx, xc);
        uint32_t a0[M9], a1[M9], a2[M9];
#pragma unroll
        for (int u = 0; u < M9; ++u) a0[u] = a1[u] = a2[u] = 0u;
        int a = 0;
        auto step = [&](auto uc) {
            constexpr int U = decltype(uc)::value;
            tc_sparse_step<U>(y, a0, a1, a2, xc);
            if (a + 1 < nX) ++fx;                             // lanes past the end of their track keep their last frame
            tc_load_limbs(X, L.plane_frames, fx, xc);
            if (a >= HALO && a < nrows) {                     // row a - 8 is complete (slot (U + 1) % 9)
                constexpr int S = (U + 1) % M9;
                const int T = (int)(a0[S] << sh3.s0) + (int)(a1[S] >> sh3.s1) + (int)(a2[S] >> sh3.s2);
                const int zr = __ldg(xn + a - HALO) + yrel - T;
                const int idx = __vimin_s32_relu(zr >> shl, NBIN + 1);
                atomicAdd(&hist[idx * 32], 1u);
            }
            ++a;
        };
#pragma unroll 1
        while (a + M9 <= nrows_w) {
            step(IC<0>{}); step(IC<1>{}); step(IC<2>{}); step(IC<3>{}); step(IC<4>{});
            step(IC<5>{}); step(IC<6>{}); step(IC<7>{}); step(IC<8>{});
        }
        const int remr = nrows_w - a;
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
