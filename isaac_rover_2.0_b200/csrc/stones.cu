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
    // eight stones per lane in flight: the loop is a chain of L2 round trips otherwise (a goal draw against 2000 stones took 19 us
    // per attempt, and the reset kernel of a 4096-env step 92 us, with one load outstanding)
    int s = lane;
    for (; s + 7 * 32 < S; s += 8 * 32) {
        float e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] = stone_edge(x, y, xn, stones + (int64_t)(s + 32 * u) * 7, mm);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            nan |= (e[u] != e[u]);
            best = fminf(best, e[u]);
        }
    }
    for (; s < S; s += 32) {
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

// ------------------------------------------------------------------------------------------------------------
// Device-side reset path: what RoverTask.pre_physics_step does for the envs whose reset_buf is set
// (rover.py:356-361) without the two host synchronisations of the reference (reset_buf.nonzero() + len(), and the
// `while reset_buf_len > 0` goal loop, rover.py:356-357,547-549):
//   reset_idx book-keeping (rover.py:451-452): reset_buf = 0, progress_buf = 0;
//   generate_goals / random_goals / check_goal_collision (rover.py:533-564): a goal on the circle of `radius` around the
//     env's initial position, re-drawn until its nearest stone edge is farther than `thr` -- per env, so the loop needs no
//     global count; the reference's quirk of re-drawing env 0 on every retry (rover.py:540) is not reproduced;
//   set_targets (rover.py:566-584): target z from the heightmap grid (get_pos_height arithmetic).
// Random numbers: Philox4x32-10, counter = (global env id lo, hi, attempt, epoch lo), key = (seed lo, seed hi ^ epoch hi);
// u = (x0 >> 8) * 2^-24, alpha = fl32(2 pi) * u -- a pure function of (seed, epoch, env, attempt): results do not depend on
// how the envs are sharded over GPUs.  oracle/reset_oracle.py restates it in numpy.
// One warp per env (lanes stride over the stones, shuffle min); warps of envs that do not reset exit at once.
// ------------------------------------------------------------------------------------------------------------
#define RESET_SMEM_STONES 2048
// warp_nearest on stones staged as (x, y, radius) triples in shared memory: the direct formula of stone_edge, the same fminf order
__device__ __forceinline__ float warp_nearest_smem(float x, float y, const float* __restrict__ st, int S, int lane) {
    float best = CUDART_INF_F;
    bool nan = false;
    for (int s = lane; s < S; s += 32) {
        const float dx = F::sub(x, st[3 * s]), dy = F::sub(y, st[3 * s + 1]);
        const float e = F::sub(__fsqrt_rn(F::add(F::mul(dx, dx), F::mul(dy, dy))), st[3 * s + 2]);
        nan |= (e != e);
        best = fminf(best, e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
    nan = __any_sync(0xffffffffu, nan);
    return nan ? CUDART_NAN_F : best;
}

__global__ void reset_targets_kernel(const int64_t* reset_in, int64_t N, int64_t env_offset, uint64_t seed, uint64_t epoch,
                                     const float* __restrict__ initial_pos, float radius, const float* __restrict__ stones, int S,
                                     float thr, int max_attempts, const float* __restrict__ hm, int H0, int H1, float hscale,
                                     float inv_hscale, float vscale, float shx, float shy, float* __restrict__ target,
                                     int64_t* __restrict__ progress, int64_t* reset_out, int32_t* __restrict__ counters,
                                     int sem) {
    // When any env of the CTA resets and the stones fit, the CTA stages (x, y, radius) of every stone in shared memory first: a
    // goal draw is then 63 shared-memory rounds instead of 63 L2 round trips (the kernel's duration is the latency of its slowest
    // warp: 92 us per 4096-env step from global memory)
    __shared__ float s_stone[RESET_SMEM_STONES * 3];
    const int lane = threadIdx.x & 31;
    const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const bool mine = n < N && reset_in[n] != 0;
    const bool staged = S <= RESET_SMEM_STONES;
    if (staged) {
        if (!__syncthreads_or(mine ? 1 : 0)) return;
        for (int i = threadIdx.x; i < S; i += blockDim.x) {
            s_stone[3 * i] = stones[(int64_t)i * 7];
            s_stone[3 * i + 1] = stones[(int64_t)i * 7 + 1];
            s_stone[3 * i + 2] = stones[(int64_t)i * 7 + 6];
        }
        __syncthreads();
    }
    if (!mine) return;
    const uint64_t gid = (uint64_t)(env_offset + n);
    const float ix = initial_pos[n * 3], iy = initial_pos[n * 3 + 1];
    float x = ix, y = iy;
    int attempt = 0;
    bool ok = false;
    while (attempt < max_attempts) {
        uint32_t r0, r1, r2, r3;
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)attempt, (uint32_t)epoch, (uint32_t)seed,
                      (uint32_t)(seed >> 32) ^ (uint32_t)(epoch >> 32), r0, r1, r2, r3);
        const float u = F::mul((float)(r0 >> 8), 5.9604644775390625e-08f);           // [0, 1), 24 bits like torch.rand
        const float alpha = F::mul(6.2831854820251465f, u);                           // 2 * math.pi * rand (rover.py:557)
        x = F::add(F::add(F::mul(radius, cosf(alpha)), 0.0f), ix);                    // rover.py:561-563
        y = F::add(F::add(F::mul(radius, sinf(alpha)), 0.0f), iy);
        ++attempt;
        const float v = staged ? warp_nearest_smem(x, y, s_stone, S, lane) : warp_nearest(x, y, stones, S, false, lane);
        if (!(v <= thr)) {                                                            // rover.py:538 (NaN counts as valid there too)
            ok = true;
            break;
        }
    }
    if (lane == 0) {
        target[n * 3] = x;
        target[n * 3 + 1] = y;
        float uu = F::sub(x, shx), vv = F::sub(y, shy);
        if (sem == RVB_SEM_TORCH_CPU) { uu = __fdiv_rn(uu, hscale); vv = __fdiv_rn(vv, hscale); }
        else { uu = F::mul(uu, inv_hscale); vv = F::mul(vv, inv_hscale); }
        const float hi = (float)(H0 - 1);
        const int i = (int)rintf(fminf(fmaxf(uu, 0.f), hi));
        const int j = min((int)rintf(fminf(fmaxf(vv, 0.f), hi)), H1 - 1);
        target[n * 3 + 2] = F::mul(hm[(int64_t)i * H1 + j], vscale);
        if (progress) progress[n] = 0;
        if (reset_out) reset_out[n] = 0;
        if (counters) {
            atomicAdd(counters + 0, 1);                 // envs reset
            atomicAdd(counters + 1, attempt);           // goals drawn
            if (!ok) atomicAdd(counters + 2, 1);        // envs that ran out of attempts (goal left at the last draw)
        }
    }
}

extern "C" int rvb_reset_targets(const int64_t* reset_in, int64_t N, int64_t env_offset, uint64_t seed, uint64_t epoch,
                                 const float* initial_pos, float radius, const float* stones, int64_t S, float thr,
                                 int32_t max_attempts, const float* heightmap, int64_t H0, int64_t H1, float hscale, float vscale,
                                 float shift_x, float shift_y, float* target, int64_t* progress, int64_t* reset_out,
                                 int32_t* counters, int sem, void* stream) {
    if (N <= 0) return RVB_OK;
    RVB_REQUIRE(reset_in && initial_pos && stones && heightmap && target, "rvb_reset_targets: null pointer");
    RVB_REQUIRE(S > 0 && S < (1 << 30) && max_attempts > 0, "rvb_reset_targets: bad S or max_attempts");
    RVB_REQUIRE(H0 > 0 && H1 > 0 && H0 < (1 << 30) && H1 < (1 << 30) && hscale > 0.f, "rvb_reset_targets: bad heightmap shape or scale");
    RVB_REQUIRE(env_offset >= 0, "rvb_reset_targets: negative env_offset");
    cudaStream_t st = as_stream(stream);
    if (counters) RVB_CUDA(cudaMemsetAsync(counters, 0, 3 * sizeof(int32_t), st));
    reset_targets_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, st>>>(reset_in, N, env_offset, seed, epoch, initial_pos, radius, stones,
                                                                         (int)S, thr, max_attempts, heightmap, (int)H0, (int)H1, hscale,
                                                                         1.0f / hscale, vscale, shift_x, shift_y, target, progress,
                                                                         reset_out, counters, sem);
    RVB_LAUNCH_CHECK();
    return RVB_OK;
}
