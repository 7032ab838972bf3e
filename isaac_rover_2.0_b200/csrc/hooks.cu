// Per-step observation epilogues the reference's authors toggled by hand, one pass over obs_buf each:
//
//   rvb_obs_hooks       rover.py:326-329 (commented out in the reference, applied here in its order when enabled):
//                         obs[:, 4:] += sqrt(0.20) * randn            -> noise_std
//                         obs[:, 4:]  = F.dropout(obs[:, 4:], p=0.1)  -> dropout_p (kept values scaled by 1/(1-p))
//                         obs         = obs - 0.02                    -> offset (every column)
//                         obs[:, remove_idx + 4] = 0                  -> zero_mask (u8 per column)
//   rvb_teacher_record  rover.py:298-300,364,374-375: one row [reset_info, a0, a1, obs...] per env of the teacher data set.
//
// Random numbers: torch's generator stream (one randn tensor per step, then one bernoulli tensor) cannot be reproduced
// per element on a sharded device path; like the reset path (stones.cu) the draws are Philox4x32-10 with
// counter = (global env id lo, hi, column, epoch lo), key = (seed lo, seed hi ^ epoch hi): x0, x1 -> Box-Muller normal
// (u1 = ((x0 >> 8) + 1) 2^-24 in (0, 1], u2 = (x1 >> 8) 2^-24), x2 -> dropout (drop iff (x2 >> 8) < round(p 2^24)) --
// a pure function of (seed, epoch, env, column), independent of how envs are sharded.  oracle/hooks_oracle.py restates it.
// HBM-streaming: 2 x 4 B per element (28.7 MB each way at 4096 envs); one thread per element, coalesced along columns.
#include "common.cuh"

__global__ void obs_hooks_kernel(float* __restrict__ obs, int64_t obs_ld, int64_t N, int C, int col0, float noise_std,
                                 uint32_t drop_thr, float keep_scale, float offset, const uint8_t* __restrict__ zero_mask,
                                 uint32_t k0, uint32_t k1, uint32_t epoch_lo, int64_t env_offset) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = blockIdx.y;
    if (c >= C) return;
    float v = obs[n * obs_ld + c];
    if (c >= col0 && (noise_std != 0.f || drop_thr != 0u)) {
        const uint64_t gid = (uint64_t)(n + env_offset);
        uint32_t x0, x1, x2, x3;
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)c, epoch_lo, k0, k1, x0, x1, x2, x3);
        if (noise_std != 0.f) {
            const float u1 = __fmul_rn((float)((x0 >> 8) + 1u), 5.9604644775390625e-8f);
            const float u2 = __fmul_rn((float)(x1 >> 8), 5.9604644775390625e-8f);
            const float r = sqrtf(__fmul_rn(-2.f, logf(u1)));
            const float z = __fmul_rn(r, cosf(__fmul_rn(6.2831854820251465f, u2)));
            v = __fadd_rn(v, __fmul_rn(noise_std, z));
        }
        if (drop_thr != 0u) v = ((x2 >> 8) < drop_thr) ? 0.f : __fmul_rn(v, keep_scale);
    }
    v = __fsub_rn(v, offset);
    if (zero_mask && zero_mask[c]) v = 0.f;
    obs[n * obs_ld + c] = v;
}

extern "C" int rvb_obs_hooks(float* obs, int64_t obs_ld, int64_t N, int64_t C, int64_t col0, float noise_std, float dropout_p,
                             float offset, const uint8_t* zero_mask, uint64_t seed, uint64_t epoch, int64_t env_offset,
                             void* stream) {
    if (N <= 0 || C <= 0) return RVB_OK;
    RVB_REQUIRE(obs, "rvb_obs_hooks: null pointer");
    RVB_REQUIRE(obs_ld >= C && C < (1 << 30) && col0 >= 0 && N <= 0x7fffffff, "rvb_obs_hooks: bad shape");
    RVB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f && noise_std >= 0.f, "rvb_obs_hooks: dropout_p must be in [0,1), noise_std >= 0");
    RVB_REQUIRE(env_offset >= 0, "rvb_obs_hooks: negative env_offset");
    const uint32_t thr = (uint32_t)llrintf(dropout_p * 16777216.0f);
    const float keep = 1.0f / (1.0f - dropout_p);
    // grid.y = env: up to 65,535 per launch
    for (int64_t n0 = 0; n0 < N; n0 += 65535) {
        const int64_t nb = (N - n0 < 65535) ? N - n0 : 65535;
        dim3 grid((unsigned)ceil_div(C, 256), (unsigned)nb);
        obs_hooks_kernel<<<grid, 256, 0, as_stream(stream)>>>(obs + n0 * obs_ld, obs_ld, nb, (int)C, (int)(col0 < C ? col0 : C), noise_std,
                                                              thr, keep, offset, zero_mask, (uint32_t)seed,
                                                              (uint32_t)(seed >> 32) ^ (uint32_t)(epoch >> 32), (uint32_t)epoch,
                                                              env_offset + n0);
    }
    RVB_LAUNCH_CHECK();
    return RVB_OK;
}

// record[n] = [reset_info[n], actions[n,0], actions[n,1], obs[n, 0..C)]   (rover.py:299,364,374-375)
__global__ void teacher_record_kernel(const float* __restrict__ reset_info, const float* __restrict__ actions, int64_t actions_ld,
                                      const float* __restrict__ obs, int64_t obs_ld, int C, float* __restrict__ record,
                                      int64_t record_ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;      // column of the record row
    const int64_t n = blockIdx.y;
    if (c >= C + 3) return;
    float v;
    if (c == 0) v = reset_info[n];
    else if (c < 3) v = actions[n * actions_ld + (c - 1)];
    else v = obs[n * obs_ld + (c - 3)];
    record[n * record_ld + c] = v;
}

extern "C" int rvb_teacher_record(const float* reset_info, const float* actions, int64_t actions_ld, const float* obs,
                                  int64_t obs_ld, int64_t N, int64_t C, float* record, int64_t record_ld, void* stream) {
    if (N <= 0) return RVB_OK;
    RVB_REQUIRE(reset_info && actions && obs && record, "rvb_teacher_record: null pointer");
    RVB_REQUIRE(C > 0 && C < (1 << 30) && obs_ld >= C && record_ld >= C + 3 && actions_ld >= 2, "rvb_teacher_record: bad shape");
    for (int64_t n0 = 0; n0 < N; n0 += 65535) {
        const int64_t nb = (N - n0 < 65535) ? N - n0 : 65535;
        dim3 grid((unsigned)ceil_div(C + 3, 256), (unsigned)nb);
        teacher_record_kernel<<<grid, 256, 0, as_stream(stream)>>>(reset_info + n0, actions + n0 * actions_ld, actions_ld,
                                                                   obs + n0 * obs_ld, obs_ld, (int)C, record + n0 * record_ld, record_ld);
    }
    RVB_LAUNCH_CHECK();
    return RVB_OK;
}
