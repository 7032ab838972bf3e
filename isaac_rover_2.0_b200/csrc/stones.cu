// stone_info validation (rover.py:533-542 check_goal_collision, :649-661 avoid_pos_rock_collision) and
// get_pos_height (rover.py:588-608).  One warp per query point, lanes stride over the S stones
// (coalesced [S,7] rows re-read from L2), warp-shuffle min.
#include <math_constants.h>

#include "common.cuh"

using F = Ops<float>;

// torch.cdist(p=2): direct formula up to 25 rows, otherwise the matmul formulation of ATen's
// _euclidean_dist: [-2x, |x|^2, 1] . [y, 1, |y|^2], clamp_min(0), sqrt.  The GEMM's accumulation order is
// not specified by torch; the 4-term dot product is evaluated here as an FMA chain in k order.
__device__ __forceinline__ float stone_edge(float x, float y, float xn, const float* __restrict__ st, bool mm) {
    const float sx = st[0], sy = st[1], r = st[6];
    float d;
    if (mm) {
        const float sn = F::add(F::mul(sx, sx), F::mul(sy, sy));
        float acc = F::mul(F::mul(x, -2.f), sx);
        acc = __fmaf_rn(F::mul(y, -2.f), sy, acc);
        acc = __fmaf_rn(xn, 1.f, acc);
        acc = __fmaf_rn(1.f, sn, acc);
        d = __fsqrt_rn(fmaxf(acc, 0.f));
    } else {
        const float dx = F::sub(x, sx), dy = F::sub(y, sy);
        d = __fsqrt_rn(F::add(F::mul(dx, dx), F::mul(dy, dy)));
    }
    return F::sub(d, r);
}

__device__ __forceinline__ float warp_nearest(float x, float y, const float* __restrict__ stones, int S, bool mm, int lane) {
    const float xn = F::add(F::mul(x, x), F::mul(y, y));
    float best = CUDART_INF_F;
    bool nan = false;
    for (int s = lane; s < S; s += 32) {
        const float e = stone_edge(x, y, xn, stones + (int64_t)s * 7, mm);
        nan |= (e != e);
        best = fminf(best, e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
    nan = __any_sync(0xffffffffu, nan);
    return nan ? CUDART_NAN_F : best;       // torch.min propagates NaN
}

__global__ void stone_validate_kernel(const float* __restrict__ xy, int64_t ld, int64_t M, const float* __restrict__ stones,
                                      int S, float thr, bool mm, float* __restrict__ nearest, int64_t* __restrict__ flag,
                                      int32_t* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    const float v = warp_nearest(xy[m * ld], xy[m * ld + 1], stones, S, mm, lane);
    if (lane == 0) {
        if (nearest) nearest[m] = v;
        const bool f = v <= thr;
        if (flag) flag[m] = f ? 1 : 0;
        if (count && f) atomicAdd(count, 1);
    }
}

extern "C" int rvb_stone_validate(const float* xy, int64_t xy_ld, int64_t M, const float* stones, int64_t S, float thr,
                                  int force_mode, float* nearest, int64_t* flag, int32_t* count, void* stream) {
    if (M <= 0) return RVB_OK;
    RVB_REQUIRE(xy && stones, "rvb_stone_validate: null pointer");
    RVB_REQUIRE(xy_ld >= 2 && S > 0 && S < (1 << 30), "rvb_stone_validate: bad xy_ld or S");
    RVB_REQUIRE(force_mode >= 0 && force_mode <= 2, "rvb_stone_validate: force_mode must be 0, 1 or 2");
    const bool mm = force_mode == 2 || (force_mode == 0 && (M > 25 || S > 25));
    stone_validate_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, as_stream(stream)>>>(xy, xy_ld, M, stones, (int)S, thr,
                                                                                        mm, nearest, flag, count);
    RVB_LAUNCH_CHECK();
    return RVB_OK;
}

// avoid_pos_rock_collision: envs are independent, so the reference's global "repeat until nothing moved"
// loop is each env's own loop; the sweep count the reference would execute is max over envs + 1.
__global__ void spawn_validate_kernel(float* __restrict__ pos, int64_t N, const float* __restrict__ stones, int S, bool mm,
                                      int max_iter, int32_t* __restrict__ iterations) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    float x = pos[n * 3], y = pos[n * 3 + 1];
    int it = 0;
    while (it < max_iter) {
        const float v = warp_nearest(x, y, stones, S, mm, lane);
        if (!(v <= 1.4f)) break;
        x = F::add(x, 0.05f);
        ++it;
    }
    if (lane == 0) {
        pos[n * 3] = x;
        if (iterations) atomicMax(iterations, it + 1);
    }
}

extern "C" int rvb_spawn_validate(float* pos, int64_t N, const float* stones, int64_t S, int32_t max_iter,
                                  int32_t* iterations, void* stream) {
    if (N <= 0) return RVB_OK;
    RVB_REQUIRE(pos && stones, "rvb_spawn_validate: null pointer");
    RVB_REQUIRE(S > 0 && S < (1 << 30) && max_iter > 0, "rvb_spawn_validate: bad S or max_iter");
    cudaStream_t st = as_stream(stream);
    if (iterations) RVB_CUDA(cudaMemsetAsync(iterations, 0, sizeof(int32_t), st));
    const bool mm = (N > 25 || S > 25);
    spawn_validate_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, st>>>(pos, N, stones, (int)S, mm, max_iter, iterations);
    RVB_LAUNCH_CHECK();
    return RVB_OK;
}

__global__ void height_lookup_kernel(const float* __restrict__ hm, int H0, int H1, const float* __restrict__ xy, int64_t ld,
                                     int64_t M, float hscale, float inv_hscale, float vscale, float shx, float shy,
                                     float* __restrict__ out, int sem) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    float u = F::sub(xy[m * ld], shx), v = F::sub(xy[m * ld + 1], shy);
    if (sem == RVB_SEM_TORCH_CPU) { u = __fdiv_rn(u, hscale); v = __fdiv_rn(v, hscale); }
    else { u = F::mul(u, inv_hscale); v = F::mul(v, inv_hscale); }
    const float hi = (float)(H0 - 1);                       // both axes clamp to size(0)-1 (rover.py:592)
    const int i = (int)rintf(fminf(fmaxf(u, 0.f), hi));
    const int j = min((int)rintf(fminf(fmaxf(v, 0.f), hi)), H1 - 1);
    out[m] = F::mul(hm[(int64_t)i * H1 + j], vscale);
}

extern "C" int rvb_height_lookup(const float* heightmap, int64_t H0, int64_t H1, const float* xy, int64_t xy_ld, int64_t M,
                                 float hscale, float vscale, float shift_x, float shift_y, float* out, int sem,
                                 void* stream) {
    if (M <= 0) return RVB_OK;
    RVB_REQUIRE(heightmap && xy && out, "rvb_height_lookup: null pointer");
    RVB_REQUIRE(H0 > 0 && H1 > 0 && H0 < (1 << 30) && H1 < (1 << 30) && xy_ld >= 2 && hscale > 0.f,
                "rvb_height_lookup: bad shape or scale");
    height_lookup_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, as_stream(stream)>>>(
        heightmap, (int)H0, (int)H1, xy, xy_ld, M, hscale, 1.0f / hscale, vscale, shift_x, shift_y, out, sem);
    RVB_LAUNCH_CHECK();
    return RVB_OK;
}
