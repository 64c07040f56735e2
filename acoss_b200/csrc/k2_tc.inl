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
//   * one CTA per SM = 128 owned windows x all streamed windows: the owned operand (15 tiles of 128 rows x 32 bytes) is
//     written once into tensor memory and every MMA takes A from there (TS form); two accumulator buffers, so the
//     tensor cores fill block b + 1 while the consumers read block b; warp roles: TMA loader, two MMA issuers, sixteen
//     consumer warps;
//   * consumer threads read their owned window's items from TMEM (tcgen05.ld 32x32b: TMEM lane = owned window,
//     column = streamed window) and bin / classify them: no dot products, no sliding sums, no shuffles on the
//     CUDA cores.
// T = acc0 * 2^e0 + ((acc1 * 256 + acc2) >> (16 - e0)) is 2 <X, Y> in the fixed-point unit of the path (e0 = 56 - 2 q_exp -
// fx_exp); the sparse level evaluates the same integer formula with dp4a, so every kernel sees identical items.
// Error budget against the exact item, in fixed-point units: quantisation <= 1.4 * 2^(e0/2), dropped limb
// products <= 0.84 * 2^e0, one floor, the norm roundings and the oracle's float32 roundings (17): 84 at e0 = 6, inside
// EPS = 128; acoss_set_tracks enables these sweeps only for e0 <= 6.

struct TcShift { int s0, s1, s2; unsigned m0; };   // e0, 8 - e0, 16 - e0, 2^e0

constexpr int TC_AF = 144;                          // owned frames held per plane (128 windows + 9 taps, rounded)
constexpr int TC_ALOAD = 137;                       // owned frames loaded per plane
constexpr int TC_BF = TC_N + 8;                     // streamed frames per plane and stage
constexpr int TC_BST = 3;                           // streamed-plane stages
constexpr int TC_PARTS = TC_N / 16;                 // 16-window parts of a block: one consumer warp per (TMEM lane quarter, part)
constexpr int TC_CONS = 4 * TC_PARTS * 32;          // consumer threads: warp w reads TMEM lanes 32 (w & 3) .., part w >> 2
constexpr int TC_THREADS = TC_CONS + 96;            // + three producer warps: two MMA issuers (the first allocates TMEM), one TMA loader
constexpr int TC_TMEM_COLS = 512;                   // two buffers of three accumulators (TC_N columns each) + the owned operand tiles
constexpr int TC_BUF = 3 * TC_N;                    // columns of one accumulator buffer
constexpr int TC_ACOL = 2 * TC_BUF;                 // first column of the owned operand: 15 tiles (plane, tap pair) of 8 columns
static_assert(TC_ACOL + 15 * 8 <= 512 && TC_N % 16 == 0, "accumulators and owned tiles must fit the tensor memory");

struct alignas(128) TcSmem {
    uint8_t a[3][TC_AF][16];                        // owned planes h, l1, l2
    uint8_t b[TC_BST][3][TC_BF][16];                // streamed planes
    uint8_t zero[TC_N][16];
    unsigned long long afull, aready, bfull[TC_BST], bfree[TC_BST], acc_full[2], acc_free[2], lvl;
    uint32_t tmem_base;
    int go, nleft;                                  // histogram sweeps: sweep again? / lines still live after a level
};

// 2 <X, Y> in fixed-point units from the three limb-product accumulators (one floor; the same expression in every kernel)
__device__ __forceinline__ int tc_item(int v0, int v1, int v2, const TcShift &sh) {
    const unsigned low = ((unsigned)v1 << 8) + (unsigned)v2;          // < 2^32: v1 <= 2 * 108 * 255^2, v2 <= 3 * 108 * 255^2
    return (int)((unsigned)v0 * sh.m0 + (low >> sh.s2));      // IMAD + SHF + IMAD: one instruction on the integer ALU pipe
}

// K-major, no swizzle: ((8, m), 2) : ((16 B, SBO), LBO); version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)(128u >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor: s32 accumulators, u8 x u8, both K-major, M = 128, N = TC_N
constexpr uint32_t TC_IDESC = (2u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// A operand from tensor memory (lane = row, 8 columns = the row's 32 bytes)
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
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

// the 30 MMAs of one block: acc0 = h.h, acc1 = h.l1 + l1.h, acc2 = l1.l1 + h.l2 + l2.h  (owned plane, streamed plane).
// The owned operand comes from tensor memory (tile (plane, tap pair) = rows' frames [m + 2 tp | m + 2 tp + 1], written once per
// CTA by tc_fill_owned): shared memory then only delivers the streamed side (2 KB per MMA instead of 6 KB; the consumers'
// shared-memory atomics and the MMA's operand fetch otherwise fight for the same 128 B / clk).
__device__ __forceinline__ void tc_issue_block(uint32_t b0, uint32_t zero, uint32_t tmem, uint32_t acc_col) {
    constexpr uint32_t PB = TC_BF * 16;
    auto prod = [&](uint32_t pa, uint32_t pb, uint32_t acc, bool first) {
        const uint32_t sb = b0 + pb * PB;
#pragma unroll
        for (int tp = 0; tp < 5; ++tp)
            tc_mma_ts(tmem + acc_col + acc * TC_N, tmem + TC_ACOL + 8 * (pa * 5 + tp),
                      tc_desc(sb + 32 * tp, tp == 4 ? zero - (sb + 32 * tp) : 16), !(first && tp == 0));   // tap 8 pairs with zeros
    };
    prod(0, 0, 0, true);
    prod(0, 1, 1, true);
    prod(1, 0, 1, false);
    prod(1, 1, 2, true);
    prod(0, 2, 2, false);
    prod(2, 0, 2, false);
}
// consumer warps 0..3 (TMEM lane quarters): owned planes (shared memory, landed on afull) -> the 15 operand tiles
__device__ __forceinline__ void tc_fill_owned(TcSmem *s, uint32_t tmem, int warp, int lane) {
    mbar_wait(&s->afull, 0);
    const int m = warp * 32 + lane;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
        for (int tp = 0; tp < 5; ++tp) {
            const uint4 f0 = *reinterpret_cast<const uint4 *>(&s->a[pl][m + 2 * tp][0]);
            const uint4 f1 = *reinterpret_cast<const uint4 *>(&s->a[pl][m + 2 * tp + 1][0]);
            const uint32_t v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
            tc_st8(tmem + ((uint32_t)(warp * 32) << 16) + TC_ACOL + 8 * (pl * 5 + tp), v);
        }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s->aready);
}

// Producer warps (one elected lane each).  Loader: owned planes once, then the streamed planes of every block (three
// stages).  Two MMA issuers, one per accumulator buffer (even / odd blocks): while one waits for its buffer or commits, the
// other keeps the tensor pipe fed (a single issuer spends a third of its time in barrier waits and copy issue).
__device__ __forceinline__ void tc_loader(TcSmem *s, const uint8_t *__restrict__ own, const uint8_t *__restrict__ str, size_t plane_bytes,
                                          int own_first, int nblocks, int max_levels) {
    mbar_expect_tx(&s->afull, 3 * TC_ALOAD * 16);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) bulk_g2s(&s->a[pl][0][0], own + pl * plane_bytes + (size_t)own_first * 16, TC_ALOAD * 16, &s->afull);
    for (int level = 0, g = 0; level < max_levels; ++level) {
        if (level > 0) {                                                  // another sweep only if the consumers ask for it
            mbar_wait(&s->lvl, (uint32_t)((level - 1) & 1));
            if (!s->go) return;
        }
        for (int b = 0; b < nblocks; ++b, ++g) {
            const int st = g % TC_BST;
            if (g >= TC_BST) mbar_wait(&s->bfree[st], (uint32_t)((g / TC_BST - 1) & 1));   // the MMAs of block g - 3 have read the stage
            mbar_expect_tx(&s->bfull[st], 3 * TC_BF * 16);
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
                bulk_g2s(&s->b[st][pl][0][0], str + pl * plane_bytes + (size_t)b * TC_N * 16, TC_BF * 16, &s->bfull[st]);
        }
    }
}
__device__ __forceinline__ void tc_issuer(TcSmem *s, int buf, int nblocks, int max_levels, uint32_t tmem) {
    mbar_wait(&s->aready, 0);
    const uint32_t zero = smem_u32(&s->zero[0][0]);
    for (int level = 0; level < max_levels; ++level) {
        if (level > 0) {
            mbar_wait(&s->lvl, (uint32_t)((level - 1) & 1));
            if (!s->go) return;
        }
        for (int g = level * nblocks; g < (level + 1) * nblocks; ++g) {
            if ((g & 1) != buf) continue;
            const int st = g % TC_BST;
            mbar_wait(&s->bfull[st], (uint32_t)((g / TC_BST) & 1));
            if (g >= 2) mbar_wait(&s->acc_free[buf], (uint32_t)(((g >> 1) - 1) & 1));   // consumers hold block g - 2 in registers
            tc_fence_after();
            tc_issue_block(smem_u32(&s->b[st][0][0][0]), zero, tmem, buf * TC_BUF);
            tc_commit(&s->acc_full[buf]);
            tc_commit(&s->bfree[st]);
        }
    }
}
// role dispatch of the three producer warps (warp index relative to the first producer warp).  max_levels > 1: after a sweep
// the producers wait for the consumers' decision (lvl barrier, s->go) whether the CTA sweeps again.
__device__ __forceinline__ void tc_producers(TcSmem *s, int pwarp, int lane, const uint8_t *__restrict__ own, const uint8_t *__restrict__ str,
                                             size_t plane_bytes, int own_first, int nblocks, uint32_t tmem, int max_levels = 1) {
    if (lane != 0) return;
    if (pwarp == 2) tc_loader(s, own, str, plane_bytes, own_first, nblocks, max_levels);
    else tc_issuer(s, pwarp, nblocks, max_levels, tmem);
}

// Common CTA prologue / epilogue: barriers, zero block, TMEM allocation by the producer warp.
__device__ __forceinline__ uint32_t tc_begin(TcSmem *s) {
    const int tid = threadIdx.x;
    for (int i = tid; i < TC_N * 16 / 4; i += TC_THREADS) reinterpret_cast<uint32_t *>(&s->zero[0][0])[i] = 0u;
    if (tid == 0) {
        mbar_init(&s->afull, 1);
        mbar_init(&s->aready, 4);
        mbar_init(&s->lvl, 1);
        s->go = 0; s->nleft = 0;
        for (int i = 0; i < TC_BST; ++i) { mbar_init(&s->bfull[i], 1); mbar_init(&s->bfree[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s->acc_full[i], 1); mbar_init(&s->acc_free[i], TC_CONS / 32); }
        mbar_fence_init();
    }
    if (tid >= TC_CONS && tid < TC_CONS + 32) {
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
    if (threadIdx.x >= TC_CONS && threadIdx.x < TC_CONS + 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
}

__device__ __forceinline__ void tc_prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Consumer side of one block: fn(first_window, acc0[16], acc1[16], acc2[16]) for this warp's 16 windows.  The accumulator
// buffer is released (acc_free) as soon as the values sit in registers.
template <typename Fn>
__device__ __forceinline__ void tc_consume_block(TcSmem *s, uint32_t tmem, int g, int b, int warp, Fn &&fn) {   // g: running block index
    const int buf = g & 1;
    mbar_wait(&s->acc_full[buf], (uint32_t)((g >> 1) & 1));
    tc_fence_after();
    const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * TC_BUF + (warp >> 2) * 16);
    int v0[16], v1[16], v2[16];
    tc_ld16(tbase, v0);
    tc_ld16(tbase + TC_N, v1);
    tc_ld16(tbase + 2 * TC_N, v2);
    tc_ld_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&s->acc_free[buf]);       // one arrival per consumer warp
    fn(b * TC_N + (warp >> 2) * 16, v0, v1, v2);
}

// ------------------------------------------------------------------------------------------------
// histogram sweep on the tensor cores.  ORIENT as in fast_hist_kernel: 0 = owned reference columns, streamed query
// windows; 1 = owned query rows, streamed reference windows.  One CTA = 128 owned lines of one pair.
// ------------------------------------------------------------------------------------------------
template <int ORIENT>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_hist_kernel(TrackSet ts, const int32_t *__restrict__ pairs, int64_t first, int n,
                                                                FastLayout L, char *__restrict__ scratch, int strips_max, TcShift sh3,
                                                                uint32_t *__restrict__ status, uint32_t *__restrict__ dbg,
                                                                uint32_t *__restrict__ dbg2, uint32_t *__restrict__ glive, uint32_t gcap) {
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
    // a line is live while its bracket can still be split (shift >= 0 in sh_a; -1 = done)
    bool valid = cons && j < My && sh_a[j] >= 0;
    if (__syncthreads_count(valid && warp < 4) == 0) return;
    auto to_sparse = [&](int line) {                          // the sparse level finishes it (call-wide list)
        const unsigned pos = atomicAdd(glive, 1u);
        if (pos < gcap) glive[1 + pos] = ((uint32_t)slot << 16) | ((uint32_t)side << 15) | (uint32_t)line;
        else atomicOr(&status[k], PAIR_ST_FALLBACK | 8u);     // reason 8: too many crowded lines in this call
    };
#ifdef TC_PHASES
    const long long T0 = clock64();
    long long Tf = 0, T3 = 0, T4 = 0;
#endif
    for (int i = tid; i < (NBIN + 2) * 128; i += TC_THREADS) hist[i] = 0u;
    const uint32_t tmem = tc_begin(s);
#ifdef TC_PHASES
    const long long T1 = clock64();
#endif
    const int nblocks = (Mxs + TC_N - 1) / TC_N;
    constexpr int LEVELS = 2;                                 // dense levels a CTA may run back to back (operand tiles stay in TMEM)
    if (!cons) {
        const uint8_t *qpl = slot_ptr<uint8_t>(scratch, L, slot, L.off_qpl), *rpl = slot_ptr<uint8_t>(scratch, L, slot, L.off_rpl);
        tc_producers(s, warp - TC_CONS / 32, lane, ORIENT == 0 ? rpl : qpl, ORIENT == 0 ? qpl : rpl, (size_t)L.plane_frames * 16, cb, nblocks, tmem,
                     LEVELS);
    } else {
        if (warp < 4) tc_fill_owned(s, tmem, warp, lane);
        uint32_t *hp = hist + tl;
        const int fk = h->fk[side], ck = h->ck[side];
        for (int level = 0; level < LEVELS; ++level) {
            const int shf = valid ? sh_a[j] : 0;
            // bin = ((z - lo) >> sh) + 1 clamped to [0, NBIN + 1]; idle lines land in the overflow bin
            const int ynrel = valid ? yn[j] - lo_a[j] + (1 << shf) : 0x40000000;
            for (int b = 0; b < nblocks; ++b) {
                if (lane == 0 && b + 1 < nblocks) tc_prefetch_l1(xn + (b + 1) * TC_N + (warp >> 2) * 16);   // next block's 16 norms (64 B)
                tc_consume_block(s, tmem, level * nblocks + b, b, warp, [&](int r0, const int (&v0)[16], const int (&v1)[16], const int (&v2)[16]) {
                    const int4 *xp = reinterpret_cast<const int4 *>(xn + r0);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int4 xv = __ldg(xp + g);
                        const int xb[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = 4 * g + u;
                            const int zr = xb[u] + ynrel - tc_item(v0[i], v1[i], v2[i], sh3);
                            const int idx = __vimin_s32_relu(zr >> shf, NBIN + 1);
                            atomicAdd(&hp[idx * 128], 1u);
                        }
                    }
                });
#ifdef TC_PHASES
                if (b == 0 && level == 0) Tf = clock64();
#endif
            }
#ifdef TC_PHASES
            if (level == 0) T3 = clock64();
#endif
            asm volatile("bar.sync 1, %0;" ::"n"(TC_CONS) : "memory");
            // one thread per line scans (all parts counted into one histogram)
            const bool scan = valid && warp < 4;
            int n_live = scan ? 1 : 0, n_miss = 0, n_left = 0;
            if (scan) {
                const Bracket br = split_bracket<128>(hp, 0, 0xffffffffu, fk, ck, lo_a[j], shf, h->lo1, h->hi1, bracket_target(nX - M9));
                if (br.bad) {                                 // cannot happen: the bins cover every item of the line
                    atomicOr(&status[k], PAIR_ST_FALLBACK | 4u);
                    sh_a[j] = -1;
                    valid = false;
                } else {
                    lo_a[j] = br.lo; w_a[j] = br.w; cb_a[j] = br.below; sh_a[j] = br.done ? -1 : br.sh;
                    n_miss = br.miss ? 1 : 0; n_left = br.done ? 0 : 1;
                    if (ORIENT == 1) {
                        int4 *rowpack = slot_ptr<int4>(scratch, L, slot, L.off_rowpack);
                        rowpack[j] = pack_row(yn[j], br.lo, br.w);
                    }
                }
            }
            uint32_t *dl = level == 0 ? dbg : dbg2;
            if (warp < 4) {
                n_live = __reduce_add_sync(0xffffffffu, n_live);
                n_miss = __reduce_add_sync(0xffffffffu, n_miss);
                n_left = __reduce_add_sync(0xffffffffu, n_left);
                if (lane == 0) {
                    if (warp == 0) atomicAdd(&dl[0], 1u);
                    atomicAdd(&dl[1], (unsigned)n_live);
                    if (n_miss) atomicAdd(&dl[2], (unsigned)n_miss);
                    if (n_left) { atomicAdd(&dl[3], (unsigned)n_left); atomicAdd(&s->nleft, n_left); }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_CONS) : "memory");
            // sweep again only when enough lines are still live to pay for it (short lines start from the whole item range and
            // need it); otherwise, and after the last level, the sparse refinement takes them as they are
            const int left = s->nleft;
            const bool again = level + 1 < LEVELS && left >= DENSE2_MIN_LIVE;
            valid = valid && sh_a[j] >= 0;                     // (all four parts of a line read what its scanning thread wrote)
#ifdef TC_PHASES
            if (level == 0) T4 = clock64();
#endif
            if (!again) {
                if (valid && warp < 4) to_sparse(j);
                if (tid == 0 && level + 1 < LEVELS) { s->go = 0; mbar_arrive(&s->lvl); }
                break;
            }
            for (int i = tid; i < (NBIN + 2) * 128; i += TC_CONS) hist[i] = 0u;
            asm volatile("bar.sync 1, %0;" ::"n"(TC_CONS) : "memory");
            if (tid == 0) { s->nleft = 0; s->go = 1; mbar_arrive(&s->lvl); }
        }
    }
    tc_end(tmem);
#ifdef TC_PHASES
    if (tid == 0 && ORIENT == 0) {                              // (orientation 1's counter block overlaps other counters)
        const long long T5 = clock64();
        atomicAdd(&dbg[20], (unsigned)((T1 - T0) >> 6)); atomicAdd(&dbg[21], (unsigned)((Tf - T1) >> 6));
        atomicAdd(&dbg[22], (unsigned)((T3 - Tf) >> 6)); atomicAdd(&dbg[23], (unsigned)((T4 - T3) >> 6));
        atomicAdd(&dbg[26], (unsigned)((T5 - T4) >> 6)); atomicAdd(&dbg[27], 1u);
        atomicAdd(&dbg[28], (unsigned)nblocks);
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// emit sweep on the tensor cores (owned = reference columns, streamed = query rows).  One CTA = 128 CRP columns = 4
// whole CRP words per row.  Per cell: the item from TMEM, the classification of fast_emit_kernel ("certainly 1" /
// "uncertain": inside a row or column bracket widened by 2 EPS, or near zero), one ballot per row = the CRP word of
// the warp's 32 columns; uncertain cells are staged per lane and compacted to the pair's pool once per 32 rows.
// ------------------------------------------------------------------------------------------------
constexpr int TC_STAGE = 32;                        // staged records per lane between two flushes (two blocks: every cell fits)

__global__ void __launch_bounds__(TC_THREADS, 1) tc_emit_kernel(TrackSet ts, const int32_t *__restrict__ pairs, int64_t first, int n,
                                                                FastLayout L, char *__restrict__ scratch, int groups, TcShift sh3,
                                                                uint32_t *__restrict__ crp_all, int words, int64_t crp_words) {
    extern __shared__ __align__(128) unsigned char tc_raw[];
    TcSmem *s = reinterpret_cast<TcSmem *>(tc_raw);
    uint2 *stage_all = reinterpret_cast<uint2 *>(tc_raw + sizeof(TcSmem));    // [consumer warp][TC_STAGE][32 lanes]
    uint32_t *words_all = reinterpret_cast<uint32_t *>(stage_all + (size_t)(TC_CONS / 32) * TC_STAGE * 32);   // [consumer warp][16]
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
        tc_producers(s, warp - TC_CONS / 32, lane, slot_ptr<uint8_t>(scratch, L, slot, L.off_rpl), slot_ptr<uint8_t>(scratch, L, slot, L.off_qpl),
                     (size_t)L.plane_frames * 16, cb, nblocks, tmem);
    } else {
        const int32_t *yn = slot_ptr<int32_t>(scratch, L, slot, L.off_bbi);
        const int4 *rowpack = slot_ptr<int4>(scratch, L, slot, L.off_rowpack);
        const int32_t *lo_c = slot_ptr<int32_t>(scratch, L, slot, L.off_lo) + L.max_rows;
        const int32_t *w_c = slot_ptr<int32_t>(scratch, L, slot, L.off_w) + L.max_rows;
        uint2 *pool = slot_ptr<uint2>(scratch, L, slot, L.off_pool);
        uint32_t *pool_ctr = slot_ptr<uint32_t>(scratch, L, slot, L.off_pcnt);
        const unsigned pool_cap = (unsigned)L.pool_cap;
        const int qd = warp & 3, part = warp >> 2;
        const int j = cb + qd * 32 + lane;                    // CRP column of this thread
        const bool valid = j < My;
        const int ynv = valid ? yn[j] : TC_HUGE;              // invalid: item huge => never in, never uncertain
        const int ycl = valid ? yn[j] - (lo_c[j] - 2 * EPS) : TC_HUGE;
        const unsigned cw1 = valid ? (unsigned)(w_c[j] + 4 * EPS - 1) : 0u;
        const unsigned jrec = (unsigned)j << 14;
        const int wcol = w_first + qd;                        // CRP word of this warp's 32 columns
        uint2 *stage = stage_all + (size_t)warp * TC_STAGE * 32 + lane;
        const uint32_t stage0 = smem_u32(stage);
        uint32_t *wrow = words_all + warp * 16;                // the 16 CRP words of this warp's part (lane 0 writes, lanes < 16 read)
        const uint32_t wrow0 = smem_u32(wrow);
        if (warp < 4) tc_fill_owned(s, tmem, warp, lane);
        uint32_t sp = stage0;                                  // next free staging entry of this lane (256 B apart)
        for (int b = 0; b < nblocks; ++b) {
            if (lane < 2 && b + 1 < nblocks) tc_prefetch_l1(rowpack + (b + 1) * TC_N + part * 16 + 8 * lane);   // next block's 16 rows (256 B)
            tc_consume_block(s, tmem, b, b, warp, [&](int r0, const int (&v0)[16], const int (&v1)[16], const int (&v2)[16]) {
                const unsigned rec0 = (unsigned)r0 | jrec;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int4 rp = __ldg(rowpack + r0 + i);  // pack_row(..)
                    const int T = tc_item(v0[i], v1[i], v2[i], sh3);
                    const int ar = rp.w + ynv - T;
                    const int ac = rp.x + ycl - T;
                    const unsigned bal = __ballot_sync(0xffffffffu, (ar & ac) < 0);
                    // lane 0: the row's word; every lane: the record (cell, ar) of an uncertain cell (row zone, column zone or
                    // near zero), predicated (a branch per cell would split the unrolled block into thousands of paths)
                    asm volatile(
                        "{\n"
                        ".reg .pred p, q;\n"
                        "setp.eq.u32 q, %8, 0;\n"
                        "@q st.shared.u32 [%9], %10;\n"
                        "setp.lt.u32 p, %1, %2;\n"
                        "setp.le.u32 q, %3, %4;\n"
                        "or.pred p, p, q;\n"
                        "setp.lt.s32 q, %1, %5;\n"
                        "or.pred p, p, q;\n"
                        "@p st.shared.v2.u32 [%0], {%6, %1};\n"
                        "@p add.u32 %0, %0, %7;\n"
                        "}\n"
                        : "+r"(sp)
                        : "r"(ar), "r"(rp.z), "r"(ac), "r"(cw1), "r"(rp.y), "r"(rec0 + (unsigned)i), "n"(256), "r"(lane),
                          "r"(wrow0 + 4u * (unsigned)i), "r"(bal)
                        : "memory");
                }
            });
            // lane l < 16 holds the word of row b * TC_N + 16 part + l
            __syncwarp();
            const int row = b * TC_N + part * 16 + lane;
            if (lane < 16 && row < Mx && wcol < words) crp[(int64_t)row * words + wcol] = wrow[lane];
            if (!(b & 1) && b + 1 < nblocks) { __syncwarp(); continue; }   // flush every second block
            const unsigned cnt = (sp - stage0) >> 8;
            sp = stage0;
            if (__any_sync(0xffffffffu, cnt != 0u)) {          // staged records -> the pair's pool: one atomic per warp
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
            __syncwarp();
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
                const int T = tc_item((int)a0[S], (int)a1[S], (int)a2[S], sh3);
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
                if (side == 0) slot_ptr<int4>(scratch, L, slot, L.off_rowpack)[j] = pack_row(ynj, br.lo, br.w);
                if (br.done) livel = false;
            }
        }
        __syncwarp();
    }
    if (lane == 0) { atomicAdd(&dbg[0], 1u); atomicAdd(&dbg[1], (unsigned)n_swept); }
    const unsigned left = __ballot_sync(0xffffffffu, livel);
    if (lane == 0 && left) atomicAdd(&dbg[2], (unsigned)__popc(left));
}
