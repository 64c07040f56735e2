// K2 fast path — placeholder until the sweep kernels land: reports "unsupported" so every pair
// takes the exact path.
#include "k2_fast.cuh"

bool k2_fast_supported(const acoss_params &, const SlotGeom &) { return false; }
size_t k2_fast_slot_bytes(const SlotGeom &) { return 0; }
int launch_k2_fast(const TrackSet &, const int32_t *, const int32_t *, int64_t, int, const acoss_params &,
                   const SlotGeom &, void *, size_t, uint32_t *, float *, float *, uint32_t *, cudaStream_t, int64_t *) {
    acoss_set_error("fast CRP path not built");
    return ACOSS_E_INVALID;
}
int k2_fast_collect_fallback(const uint32_t *, int64_t, int, int32_t *, int32_t *, int *count_host, cudaStream_t) {
    *count_host = 0;
    return ACOSS_OK;
}
