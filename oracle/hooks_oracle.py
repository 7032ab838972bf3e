"""TEST INFRASTRUCTURE ONLY -- numpy restatement of rvb_obs_hooks / rvb_teacher_record (include/rover_b200.h).

The arithmetic follows the reference's (commented-out) lines rover.py:326-329 in their order and the recorder's row layout
(rover.py:299-300,364,374-375).  The reference draws from torch's global generator (`torch.randn`, `F.dropout`), whose
stream cannot be reproduced per element on a sharded device path; as for the reset path the draws are Philox4x32-10
(pinned to the Random123 known answers in oracle/reset_oracle.py), so parity here is: PINNED for the generator and for the
deterministic part (offset, mask, recorder rows: tests/test_hooks_cpu.py runs the reference's own statements on the same
tensors), distributional for noise and dropout (mean / variance / drop rate / scaling).  Never imported by the product.
"""
import numpy as np

from reset_oracle import MASK, philox4x32_10

TWO_PI_F32 = np.float32(6.2831854820251465)
INV24 = np.float32(2.0 ** -24)


def draws(seed, epoch, gid, cols):
    """Philox words of element (env gid, column c): arrays [len(gid), len(cols)] x 4."""
    gid = np.asarray(gid, dtype=np.uint64)[:, None]
    c = np.asarray(cols, dtype=np.uint64)[None, :]
    g_lo, g_hi = np.broadcast_arrays(gid & MASK, gid >> np.uint64(32))
    shape = np.broadcast(gid, c).shape
    k0 = np.uint64(seed & MASK)
    k1 = np.uint64(((seed >> 32) ^ (epoch >> 32)) & MASK)
    return philox4x32_10((np.broadcast_to(g_lo, shape), np.broadcast_to(g_hi, shape), np.broadcast_to(c, shape),
                          np.full(shape, epoch & MASK, dtype=np.uint64)), (k0, k1))


def normal_from_words(x0, x1):
    f = np.float32
    u1 = ((x0 >> np.uint32(8)).astype(np.uint32) + np.uint32(1)).astype(f) * INV24
    u2 = (x1 >> np.uint32(8)).astype(f) * INV24
    r = np.sqrt((f(-2.0) * np.log(u1).astype(f)).astype(f)).astype(f)
    return (r * np.cos((TWO_PI_F32 * u2).astype(f)).astype(f)).astype(f)


def obs_hooks(obs, col0, noise_std, dropout_p, offset, zero_mask, seed, epoch, env_offset=0):
    """obs f32 [N, C] -> new array (the kernel works in place)."""
    f = np.float32
    obs = np.array(obs, dtype=f, copy=True)
    N, C = obs.shape
    if (noise_std != 0 or dropout_p != 0) and col0 < C:
        cols = np.arange(col0, C)
        x0, x1, x2, _ = draws(seed, epoch, np.arange(N) + env_offset, cols)
        v = obs[:, col0:]
        if noise_std != 0:
            v = (v + (f(noise_std) * normal_from_words(x0, x1)).astype(f)).astype(f)            # rover.py:326
        thr = np.uint32(int(np.rint(f(dropout_p) * f(16777216.0))))
        if thr != 0:
            keep = f(1.0) / (f(1.0) - f(dropout_p))
            v = np.where((x2 >> np.uint32(8)) < thr, f(0), (v * keep).astype(f)).astype(f)       # rover.py:327
        obs[:, col0:] = v
    obs = (obs - f(offset)).astype(f)                                                            # rover.py:328
    if zero_mask is not None:
        obs[:, np.asarray(zero_mask).astype(bool)] = 0                                           # rover.py:329
    return obs


def teacher_row(reset_info, actions, obs):
    """rover.py:299,364,374-375: data_curr_timestep = [reset_info | actions | obs_buf]."""
    return np.concatenate([np.asarray(reset_info, np.float32)[:, None], np.asarray(actions, np.float32)[:, :2],
                           np.asarray(obs, np.float32)], axis=1)
