// Fused env step: everything RLTask.post_physics_step (rl_task.py:239-259) and the action half of
// RoverTask.pre_physics_step (rover.py:366-414) compute for one step, enqueued by ONE host call:
//   step_pre_kernel      progress += 1, tensor_quat_to_eul, Memory.input_state x2, Ackermann + joint targets,
//                        proprioceptive observation columns + heading               (one thread per env)
//   heightmap ray-cast   Camera.get_depths with the sparse/dense gather fused into obs_buf[:, 4:]
//   rock collision       Rock_Detection.get_collisions + check_collision, on a forked stream (independent of the ray-cast)
//   reward_reset         calculate_metrics + is_done + episode statistics (rvb_reward_reset)
// The per-call entry points of rover_b200.h stay the reference-shaped API; this one removes ~25 Python/ctypes round trips
// per step (the host side was the bottleneck once the ray-cast dropped below 1.5 ms).
#include "task_dev.cuh"
#include "raycast_common.cuh"

__global__ void step_pre_kernel(rvb_step_io io, int64_t N, int H, int sem) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (io.progress) io.progress[n] += 1;                                      // rl_task.py:241
    const float4 q = reinterpret_cast<const float4*>(io.quat)[n];
    float r, p, y;
    quat_to_euler_dev(q.x, q.y, q.z, q.w, r, p, y);
    io.euler[n * 3 + 0] = r; io.euler[n * 3 + 1] = p; io.euler[n * 3 + 2] = y;
    const float lin = io.actions[n * 2 + 0], ang = io.actions[n * 2 + 1];
    float* hl = io.lin_hist + n * H;
    float* ha = io.ang_hist + n * H;
    for (int i = H - 1; i > 0; --i) { hl[i] = hl[i - 1]; ha[i] = ha[i - 1]; }  // rover.py:76-77
    hl[0] = lin; ha[0] = ang;
    float st[6], ve[6];
    ackermann_dev(lin, ang, sem, st, ve);
    ackermann_store(n, st, ve, io.steer, io.vel, io.pos_targets, io.vel_targets);
    obs_proprio_dev(io.pos[n * 3 + 0], io.pos[n * 3 + 1], y, io.target[n * 3 + 0], io.target[n * 3 + 1], lin, ang, sem,
                    io.obs + n * io.obs_ld, io.heading + n);
}

int launch_reward_reset(const rvb_reward_params* p, const float* pos, const float* target, const float* heading,
                        const float* rover_rot, const float* lin, const float* lin_prev, const float* ang,
                        const float* ang_prev, int64_t hs, const float* joints, const int64_t* progress,
                        const int64_t* rock_collision, int64_t N, float* rew, int64_t* reset, float* ex_pos_reward,
                        int64_t* ex_collision, float* ex_uprightness, float* ex_heading, float* ex_motion,
                        float* ex_goal_angle, double* stats, double* stats_scratch, cudaStream_t st);

#include <vector>

namespace {
// optional device timing of the heightmap ray-cast inside rvb_env_step (bench.py's roofline figure)
struct Timing {
    bool on = false;
    std::vector<cudaEvent_t> ev;     // start, end, start, end, ...
    size_t used = 0;
};
thread_local Timing g_timing;

}  // namespace

extern "C" int rvb_env_step(const rvb_terrain* terrain, const rvb_terrain* rocks, const rvb_reward_params* p,
                            const rvb_step_io* io_, const double* pattern, int64_t P, const int32_t* col_a,
                            const int32_t* col_b, int64_t N, int64_t H, void* stream) {
    RVB_REQUIRE(N >= 0, "rvb_env_step: N < 0");
    if (N == 0) return RVB_OK;
    RVB_REQUIRE(terrain && p && io_ && pattern && col_a && col_b, "rvb_env_step: null pointer");
    const rvb_step_io io = *io_;
    RVB_REQUIRE(io.pos && io.quat && io.joints && io.actions && io.target && io.lin_hist && io.ang_hist && io.euler && io.heading &&
                    io.obs && io.dist && io.rew && io.reset && io.progress, "rvb_env_step: null pointer in rvb_step_io");
    RVB_REQUIRE(((uintptr_t)io.quat & 15) == 0, "rvb_env_step: quat must be 16-byte aligned");
    RVB_REQUIRE(H >= 2 && H <= 64 && io.obs_ld >= 4, "rvb_env_step: bad history length or obs_ld");
    RVB_REQUIRE(!io.obs_h16 || io.obs_h16_ld > 0, "rvb_env_step: obs_h16 needs obs_h16_ld");
    RVB_REQUIRE(p->curriculum_level < 2 || (rocks && io.wheel_dist && io.body_dist && io.rock_collision),
                "rvb_env_step: curriculum_level >= 2 needs the rock layer and its outputs");
    cudaStream_t st = as_stream(stream);
    step_pre_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(io, N, (int)H, p->sem);
    RVB_LAUNCH_CHECK();
    const bool with_rocks = rocks && io.wheel_dist && io.body_dist;
    RvbJoinGuard guard(st);          // on every exit path below: st waits for the rock layer's kernel
    if (with_rocks) {
        RvbSide* side = nullptr;
        RVB_CUDA(rvb_side_stream(0, &side));
        RVB_CUDA(cudaEventRecord(side->fork, st));
        RVB_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
        const int rc = rvb_rock_collision(rocks, io.pos, io.euler, nullptr, io.joints, N, io.wheel_dist, io.body_dist, nullptr,
                                          io.rock_collision, nullptr, 0, side->s);
        if (rc != RVB_OK) return rc;
        RVB_CUDA(cudaEventRecord(side->join, side->s));
        guard.join = side->join;
    }
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (g_timing.on) {
        if (g_timing.used + 2 > g_timing.ev.size()) {
            cudaEvent_t a, b;
            RVB_CUDA(cudaEventCreate(&a));
            RVB_CUDA(cudaEventCreate(&b));
            g_timing.ev.push_back(a);
            g_timing.ev.push_back(b);
        }
        t0 = g_timing.ev[g_timing.used];
        t1 = g_timing.ev[g_timing.used + 1];
        g_timing.used += 2;
        RVB_CUDA(cudaEventRecord(t0, st));
    }
    // packed observation requested: the heightmap columns go to obs_h16 (fp16, what they are by construction) and the f32
    // columns 4.. of obs are left alone
    int rc = rvb_heightmap_raycast2(terrain, io.pos, io.euler, nullptr, pattern, P, N, io.dist, nullptr, nullptr, nullptr, nullptr,
                                    io.obs_h16 ? nullptr : io.obs, io.obs_ld, io.obs_h16, io.obs_h16_ld, 4, col_a, col_b, 0, st);
    if (rc != RVB_OK) return rc;
    if (t1) RVB_CUDA(cudaEventRecord(t1, st));
    if (guard.join) {
        RVB_CUDA(cudaStreamWaitEvent(st, guard.join, 0));
        guard.join = nullptr;
    }
    // hist[:, 0] = this step's action, hist[:, 1] = the previous one (rover.py:498-501); rover_rot = euler (no physics in between)
    return launch_reward_reset(p, io.pos, io.target, io.heading, io.euler, io.lin_hist, io.lin_hist + 1, io.ang_hist, io.ang_hist + 1, H,
                            io.joints, io.progress, p->curriculum_level >= 2 ? io.rock_collision : nullptr, N, io.rew, io.reset,
                            io.ex_pos_reward, io.ex_collision, io.ex_uprightness, io.ex_heading, io.ex_motion, io.ex_goal_angle,
                            io.stats, io.stats_scratch, st);
}

extern "C" int rvb_timing_enable(int on) {
    g_timing.on = on != 0;
    g_timing.used = 0;
    return RVB_OK;
}

extern "C" int rvb_timing_read(float* ms_host, int cap) {
    int n = 0;
    for (size_t i = 0; i + 1 < g_timing.used && n < cap; i += 2, ++n) {
        if (cudaEventSynchronize(g_timing.ev[i + 1]) != cudaSuccess) return rvb_set_error(RVB_ERR_CUDA, "rvb_timing_read", "event sync failed");
        if (cudaEventElapsedTime(ms_host + n, g_timing.ev[i], g_timing.ev[i + 1]) != cudaSuccess)
            return rvb_set_error(RVB_ERR_CUDA, "rvb_timing_read", "elapsed time failed");
    }
    g_timing.used = 0;
    return n;
}
