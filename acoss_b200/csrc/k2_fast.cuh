// K2 fast path interface (k2_fast.cu).
#pragma once
#include "common.cuh"

// true when the sweep kernels can handle these parameters (tau == 1, m <= 16, sizes in range)
bool k2_fast_supported(const acoss_params &p, const SlotGeom &g, const TrackSet &ts);
// scratch bytes per slot for the fast path
size_t k2_fast_slot_bytes(const SlotGeom &g, int max_frames);
// CRP bits + exact thresholds for pairs[first..first+n); pairs that fail a consistency check get
// PAIR_ST_FALLBACK in status[k] and are re-run by the exact path
int launch_k2_fast(const TrackSet &ts, const int32_t *pairs, const int32_t *oti, int64_t first, int n,
                   const acoss_params &p, const SlotGeom &g, void *scratch, size_t slot_bytes, uint32_t *crp,
                   float *thr_q, float *thr_r, uint32_t *status, uint32_t *dbg, cudaStream_t st, int64_t *launches,
                   KernelTimer *timer,              // optional (may be NULL): brackets every kernel with CUDA events
                   uint32_t *glive, uint32_t gcap); // call-wide list of lines for the sparse level: [0] count, [1..gcap] entries
// entries the call-wide sparse list should hold for n slots
inline uint32_t k2_fast_sparse_cap(int64_t slots) { return (uint32_t)(slots * 48 + 4096); }
// compacts the absolute indices k in [first, first+n) whose status has PAIR_ST_FALLBACK into map_dev (count in
// count_dev); with count_host != NULL it also synchronises the stream to return the count
int k2_fast_collect_fallback(const uint32_t *status, int64_t first, int n, int32_t *map_dev, int32_t *count_dev,
                             int *count_host, cudaStream_t st);
