// Offline K-nearest-centroid index builder (rover_utils.py:52-118) -- device implementation.
#include "common.cuh"

extern "C" int rvb_build_knn_index(const int32_t* triangles, int64_t T, const uint16_t* vertices, int64_t V, int64_t G0,
                                   int64_t G1, float res, int64_t K, int32_t* out, void* stream) {
    (void)triangles; (void)T; (void)vertices; (void)V; (void)G0; (void)G1; (void)res; (void)K; (void)out; (void)stream;
    return rvb_set_error(RVB_ERR_UNSUPPORTED, "rvb_build_knn_index", "not implemented yet");
}
