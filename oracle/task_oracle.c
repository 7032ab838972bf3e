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
