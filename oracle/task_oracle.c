/* TEST INFRASTRUCTURE ONLY -- never linked, loaded or called by the product.
 *
 * Plain-C restatement (fp32, torch-CPU semantics) of the kinematics on the hot path:
 *   Ackermann                 tasks/utils/kinematics.py:14-67
 *   joint-target permutation  tasks/rover.py:400-409
 * Everything but atan2 is IEEE arithmetic and must match the reference bit for bit; atan2f (libm here, Sleef in torch) may
 * differ in the last ulp.  Pinned by tests/test_c_oracle_cpu.py against the golden vectors the reference produced.
 */
#include <math.h>
#include <stdint.h>

static const float WX[6] = {-0.385f, 0.385f, -0.447f, 0.447f, -0.385f, 0.385f};      /* kinematics.py:20-25, x right */
static const float WY[6] = {0.438f, 0.438f, 0.0f, 0.0f, -0.411f, -0.411f};           /* y forward; FL FR ML MR RL RR */
static const float SIDE[6] = {-1.0f, 1.0f, -1.0f, 1.0f, -1.0f, 1.0f};                /* kinematics.py:46 */

/* lin, ang f32 [n] -> steering_angles f32 [n,6], motor_velocities f32 [n,6] */
void rvo_ackermann(const float* lin_vel, const float* ang_vel, int64_t n, float* steer, float* vel) {
    const float lo = (float)(-3.14 / 2), hi = (float)(3.14 / 2), pi = (float)3.141592653589793;       /* kinematics.py:64-65 */
    for (int64_t i = 0; i < n; ++i) {
        float lin = lin_vel[i];
        const float ang = ang_vel[i];
        float P = copysignf(lin / ang, -ang);                /* :34-35 */
        P = fabsf(P) > 0.45f ? P : 0.0f;                     /* :38  (NaN compares false -> 0) */
        lin = (P != 0.0f) ? lin : 0.0f;                      /* :39 */
        for (int w = 0; w < 6; ++w) {
            const float dx = P - WX[w], dy = 0.0f - WY[w];
            const float dist = sqrtf(dx * dx + dy * dy);     /* :42-43 */
            const float omega = (lin != 0.0f) ? copysignf(ang, lin) : ang * SIDE[w];     /* :49-52 */
            float v = dist * omega;                          /* :55 */
            if (dist > 1000.0f) v = lin;                     /* :58 */
            vel[6 * i + w] = v / 0.2f;                       /* :61 */
            float s = atan2f(WY[w], WX[w] - P);              /* :63 (both where-branches are the same expression) */
            if (s < lo) s = s + pi;                          /* :64 */
            if (s > hi) s = s - pi;                          /* :65 */
            steer[6 * i + w] = s;
        }
    }
}

/* rover.py:400-409: positions = steer[:, (FR, RR, FL, RL)], velocities = vel[:, (FR, CR, RR, FL, CL, RL)] */
void rvo_joint_targets(const float* steer, const float* vel, int64_t n, float* positions, float* velocities) {
    static const int PI_[4] = {1, 5, 0, 4}, VI_[6] = {1, 3, 5, 0, 2, 4};
    for (int64_t i = 0; i < n; ++i) {
        for (int j = 0; j < 4; ++j) positions[4 * i + j] = steer[6 * i + PI_[j]];
        for (int j = 0; j < 6; ++j) velocities[6 * i + j] = vel[6 * i + VI_[j]];
    }
}

/* RoverTask.calculate_metrics (rover.py:460-531) and is_done (rover.py:610-647), fp32 with torch's promotion rules (a Python
 * float meeting a float32 tensor is rounded to float32 first; int64 tensors times a Python float give float32).  Reward scales:
 * cfg/task/Rover.yaml:37-46.  Pure IEEE arithmetic (+ sqrt) => bit for bit.
 * extras [n,5] = pos_reward, heading_contraint_penalty, motion_contraint_penalty, goal_angle_penalty, uprightness_penalty. */
void rvo_reward_reset(const float* pos, const float* target, const float* heading, const float* lin, const float* lin_prev,
                      const float* ang, const float* ang_prev, const float* joints, int64_t joints_ld, const int64_t* progress,
                      const int64_t* rock_collision, const float* rover_rot, int level, int64_t max_len, int64_t n, float* rew,
                      int64_t* reset, float* extras) {
    const float heading_scale = 0.05f, motion_scale = -0.01f, goal_scale = 0.3f, boogie_scale = 0.5f, pos_scale = 1.0f;
    const float k33 = (float)(0.33 * 0.33), near = (float)0.18, tilt = (float)(0.78 * 1.5);
    for (int64_t i = 0; i < n; ++i) {
        const float dx = target[3 * i] - pos[3 * i], dy = target[3 * i + 1] - pos[3 * i + 1];
        const float td = sqrtf(dx * dx + dy * dy);                                         /* :482 */
        const float heading_pen = (lin[i] < 0.0f ? -1.0f : 0.0f) * heading_scale;          /* :486 */
        const float* j = joints + joints_ld * i;
        const float boogie = ((fabsf(j[0]) + fabsf(j[1])) + fabsf(j[2])) * boogie_scale;   /* :492 */
        const float hd = heading[i];
        const float goal_pen = fabsf(hd) > 2.0f ? -fabsf((hd * 0.3f) * goal_scale) : 0.0f; /* :495 */
        const float dl = fabsf(lin[i] * 3.0f - 3.0f * lin_prev[i]);                        /* :498 */
        const float da = fabsf(ang[i] * 3.0f - 3.0f * ang_prev[i]);
        const float p1 = dl > 0.05f ? dl * dl : 0.0f, p2 = da > 0.05f ? da * da : 0.0f;    /* :499-500 */
        const float motion = (p1 * p1) * motion_scale + (p2 * p2) * motion_scale;          /* :501-502 */
        float pos_rew = (1.0f / (1.0f + ((k33 * td) * td))) * pos_scale;                   /* :505 */
        if (td <= near) pos_rew = 1.03f * (float)(max_len - progress[i]);                  /* :506 */
        float r = ((pos_rew + heading_pen) + motion) + goal_pen;                           /* :512 */
        const int hit = level >= 2 && rock_collision && rock_collision[i] == 1;
        if (hit) r = r - 300.0f;                                                           /* :519 */
        rew[i] = r / 3000.0f;                                                              /* :522 */
        float* e = extras + 5 * i;
        e[0] = pos_rew; e[1] = heading_pen; e[2] = motion; e[3] = goal_pen; e[4] = boogie;
        int64_t done = progress[i] >= max_len;                                             /* :614 */
        if (fabsf(rover_rot[3 * i]) >= tilt) done = 1;                                     /* :615 */
        if (fabsf(rover_rot[3 * i + 1]) >= tilt) done = 1;                                 /* :616 */
        if (td >= 11.0f) done = 1;                                                         /* :617 */
        if (td <= near) done = 1;                                                          /* :618 */
        if (hit) done = 1;                                                                 /* :619 */
        reset[i] = done;
    }
}

/* tensor_quat_to_eul (tasks/utils/math/tensor_quat_to_euler.py:6-31), wxyz -> roll, pitch, yaw in fp32.  atan2 / asin are libm's
 * here and Sleef's in torch: equal to an ulp, not bit for bit.  torch.pi of the reference = 3.1415927410125732 (:4). */
void rvo_quat_to_euler(const float* q, int64_t n, float* e) {
    const float pi = 3.1415927410125732f;
    for (int64_t i = 0; i < n; ++i) {
        const float w = q[4 * i], x = q[4 * i + 1], y = q[4 * i + 2], z = q[4 * i + 3];
        e[3 * i] = atan2f(2.0f * (w * x + y * z), 1.0f - (2.0f * (x * x + y * y)));           /* :17-19 */
        const float sinp = 2.0f * (w * y - z * x);                                            /* :22 */
        const float t = sinp - 1.0f;
        const float sgn = (t > 0.0f) - (t < 0.0f);                                            /* torch.sign; NaN -> 0 here, NaN >= 0 false there: both take asin */
        e[3 * i + 1] = (t == t && sgn >= 0.0f) ? copysignf((1.0f * pi) / 2.0f, sinp) : asinf(sinp);   /* :23-24 */
        e[3 * i + 2] = atan2f(2.0f * (w * z + x * y), 1.0f - (2.0f * (y * y + z * z)));       /* :27-29 */
    }
}

/* The four proprioceptive observation columns and heading_diff (rover.py:279-283,320-323). */
void rvo_obs_proprio(const float* pos, const float* euler, const float* target, const float* lin_now, const float* ang_now,
                     int64_t n, float* obs4, float* heading) {
    for (int64_t i = 0; i < n; ++i) {
        const float yaw = euler[3 * i + 2];
        const float dx = cosf(yaw), dy = sinf(yaw);                                           /* :279-280 */
        const float tx = target[3 * i] - pos[3 * i], ty = target[3 * i + 1] - pos[3 * i + 1]; /* :281 */
        const float h = -atan2f(tx * dy - ty * dx, tx * dx + ty * dy);                        /* :283 */
        heading[i] = h;
        obs4[4 * i] = sqrtf(tx * tx + ty * ty) / 9.0f;                                        /* :320 */
        obs4[4 * i + 1] = h / (float)3.141592653589793;                                       /* :321 math.pi */
        obs4[4 * i + 2] = lin_now[i];                                                         /* :322 */
        obs4[4 * i + 3] = ang_now[i];                                                         /* :323 */
    }
}

/* obs_buf[:, 4:] (rover.py:324-325): sparse / 2 and dense / 2 are fp16 divisions (one rounding), stored as fp32.
 * dist = fp16 bits [P] of one env; idx = the pattern's coarse then fine index vectors (heightmap_distribution.py:107-133). */
void rvo_obs_heightmap(const uint16_t* dist, const int64_t* idx, int64_t n_idx, float* out) {
    for (int64_t i = 0; i < n_idx; ++i) {
        _Float16 h;
        __builtin_memcpy(&h, dist + idx[i], 2);
        out[i] = (float)(_Float16)((float)h / 2.0f);
    }
}

/* get_pos_height (rover.py:588-608): idx = round_half_even(clamp((xy - shift) / hscale, 0, H - 1)); heightmap[ix, iy] * vscale */
void rvo_pos_height(const float* heightmap, int64_t H0, int64_t H1, const float* xy, int64_t xy_ld, int64_t n, float hscale,
                    float vscale, float shift_x, float shift_y, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        float u = (xy[xy_ld * i] - shift_x) / hscale, v = (xy[xy_ld * i + 1] - shift_y) / hscale;
        u = fminf(fmaxf(u, 0.0f), (float)(H0 - 1));                                           /* the reference clamps both to size()[0] - 1 */
        v = fminf(fmaxf(v, 0.0f), (float)(H0 - 1));
        out[i] = heightmap[(int64_t)rintf(u) * H1 + (int64_t)rintf(v)] * vscale;
    }
}

/* Stones (terrain_utils.py:416-424, rover.py:533-542,649-661).
 * read_stone_info: column 6 = max(col 3, col 4) / 4 in the file's dtype (f64), then the whole row cast to f32.
 * nearest stone edge = min_s(||xy - stone_s.xy|| - radius_s): the direct formula torch.cdist uses up to 25 rows (bit for bit);
 * above that torch switches to its matmul formulation (|a|^2 + |b|^2 - 2ab), which agrees to ~1e-4 only.
 * check_goal_collision: invalid iff nearest <= 1.0 (rover.py:539); avoid_pos_rock_collision: x += 0.05 while nearest <= 1.4
 * for any env, until nothing changes (rover.py:655-660). */
void rvo_read_stone_info(const double* stone6, int64_t S, float* stone7) {
    for (int64_t s = 0; s < S; ++s) {
        const double* r = stone6 + 6 * s;
        for (int c = 0; c < 6; ++c) stone7[7 * s + c] = (float)r[c];
        stone7[7 * s + 6] = (float)((r[3] > r[4] ? r[3] : r[4]) / 4);
    }
}

void rvo_nearest_stone_edge(const float* xy, int64_t xy_ld, int64_t n, const float* stone7, int64_t S, float* nearest) {
    for (int64_t i = 0; i < n; ++i) {
        float best = INFINITY;
        for (int64_t s = 0; s < S; ++s) {
            const float dx = xy[xy_ld * i] - stone7[7 * s], dy = xy[xy_ld * i + 1] - stone7[7 * s + 1];
            const float d = sqrtf(dx * dx + dy * dy) - stone7[7 * s + 6];
            if (d < best) best = d;
        }
        nearest[i] = best;
    }
}

/* pos f32 [n,3] in place; returns the number of sweeps */
int64_t rvo_avoid_pos_rock_collision(float* pos, int64_t n, const float* stone7, int64_t S, int64_t max_iter) {
    for (int64_t it = 1; it <= max_iter; ++it) {
        int changed = 0;
        for (int64_t i = 0; i < n; ++i) {
            float near;
            rvo_nearest_stone_edge(pos + 3 * i, 3, 1, stone7, S, &near);
            if (near <= 1.4f) { pos[3 * i] = pos[3 * i] + 0.05f; changed = 1; }
        }
        if (!changed) return it;
    }
    return -1;
}
