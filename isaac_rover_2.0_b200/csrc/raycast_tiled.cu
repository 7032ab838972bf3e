// Production heightmap ray-cast kernel (placeholder: routes to the simple kernel in env chunks).
#include "common.cuh"

int launch_heightmap_tiled(const rvb_terrain* t, const float* pos, const float* euler, const float* trig,
                           const double* pattern, int64_t P, int64_t N, uint16_t* dist, int32_t* hit_slot,
                           int32_t* hit_tri, uint16_t* pt, uint16_t* sources, float* obs, int64_t obs_ld,
                           const int32_t* col_a, const int32_t* col_b, cudaStream_t st) {
    const int64_t chunk = 32768;
    for (int64_t n0 = 0; n0 < N; n0 += chunk) {
        const int64_t n = (N - n0 < chunk) ? N - n0 : chunk;
        int rc = rvb_heightmap_raycast(t, pos + n0 * 3, euler + n0 * 3, trig ? trig + n0 * 6 : nullptr, pattern, P, n,
                                       dist + n0 * P, hit_slot ? hit_slot + n0 * P : nullptr,
                                       hit_tri ? hit_tri + n0 * P : nullptr, pt ? pt + n0 * P * 3 : nullptr,
                                       sources ? sources + n0 * P * 3 : nullptr, obs ? obs + n0 * obs_ld : nullptr, obs_ld,
                                       col_a, col_b, 1, (void*)st);
        if (rc != RVB_OK) return rc;
    }
    return RVB_OK;
}
